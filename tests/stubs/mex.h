/* Declarations-only stand-in for MATLAB's mex.h (see matrix.h in this directory). */
#ifndef HPRLP_TEST_MEX_STUB_H
#define HPRLP_TEST_MEX_STUB_H
#include "matrix.h"
#ifdef __cplusplus
extern "C" {
#endif
void mexErrMsgIdAndTxt(const char *, const char *, ...);
void mexWarnMsgIdAndTxt(const char *, const char *, ...);
int mexPrintf(const char *, ...);
int mexEvalString(const char *);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
