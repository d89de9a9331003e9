/* Declarations-only stand-in for MATLAB's C Matrix API (matrix.h), just enough to COMPILE the reference's MEX gateway in
 * tests/test_dropin_sources.py (no MATLAB in the image; nothing here is linked or run).  Signatures as documented by
 * MathWorks for the interleaved/separate-complex C API. */
#ifndef HPRLP_TEST_MATRIX_STUB_H
#define HPRLP_TEST_MATRIX_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS, mxDOUBLE_CLASS,
               mxSINGLE_CLASS, mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS, mxINT32_CLASS, mxUINT32_CLASS,
               mxINT64_CLASS, mxUINT64_CLASS, mxFUNCTION_CLASS } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
double *mxGetPr(const mxArray *);
void *mxGetData(const mxArray *);
double mxGetScalar(const mxArray *);
size_t mxGetM(const mxArray *);
size_t mxGetN(const mxArray *);
size_t mxGetNumberOfElements(const mxArray *);
mwSize mxGetNzmax(const mxArray *);
mwIndex *mxGetJc(const mxArray *);
mwIndex *mxGetIr(const mxArray *);
mxArray *mxGetField(const mxArray *, mwIndex, const char *);
void mxSetField(mxArray *, mwIndex, const char *, mxArray *);
void mxSetCell(mxArray *, mwIndex, mxArray *);
mxArray *mxCreateDoubleScalar(double);
mxArray *mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray *mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray *mxCreateStructMatrix(mwSize, mwSize, int, const char **);
mxArray *mxCreateCellMatrix(mwSize, mwSize);
mxArray *mxCreateString(const char *);
char *mxArrayToString(const mxArray *);
void mxFree(void *);
bool mxIsDouble(const mxArray *);
bool mxIsUint64(const mxArray *);
bool mxIsUint32(const mxArray *);
bool mxIsEmpty(const mxArray *);
bool mxIsComplex(const mxArray *);
bool mxIsSparse(const mxArray *);
bool mxIsChar(const mxArray *);
bool mxIsLogicalScalar(const mxArray *);
bool mxIsLogicalScalarTrue(const mxArray *);
bool mxIsStruct(const mxArray *);
bool mxIsCell(const mxArray *);
bool mxIsNumeric(const mxArray *);
bool mxIsLogical(const mxArray *);
#ifdef __cplusplus
}
#endif
#endif
