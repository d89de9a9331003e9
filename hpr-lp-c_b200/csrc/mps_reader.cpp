// mps_reader.cpp -- host-side free-format MPS reader (stays on the host; OUT OF SCOPE as a
// rebuild target, needed so create_model_from_mps / build/solve_mps_file work unchanged).
// Behaviour mirrors the reference reader (src/mps_reader.cpp) so the resulting CSR / bound /
// cost arrays are bit-identical:
//   * files are read as FREE format only (src/mps_reader.cpp:1517); '*' and '&' lines are comments;
//   * first N row is the objective, later N rows are ignored ("rim") (:598-612);
//   * RHS on the objective row sets obj_constant = -value (:765-767);
//   * RANGES: E rows take the sign of the range, L/G rows use |range| (:808-836);
//   * bounds default to [0,inf), [0,1] for integer-marked columns; an upper bound < 0 with no
//     lower bound gives l = -inf (:1150-1180); OBJSENSE is parsed and ignored (quirk 12);
//   * COO -> CSR sorts by (row, col) and sums duplicates (:1266-1361); numbers via atof.
// Written from scratch with STL containers (the reference uses a djb2 hash + qsort in C style).
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

#include "engine.h"

namespace hpr {
namespace {

enum Section { S_NONE, S_OBJSENSE, S_ROWS, S_COLUMNS, S_RHS, S_BOUNDS, S_RANGES, S_OTHER };

struct Coo { int row, col; double val; };

struct MpsData {
    std::unordered_map<std::string, int> row_index;   // 0 objective, -1 rim objective, k+1 constraint k
    std::unordered_map<std::string, int> col_index;
    std::vector<char> row_type;                        // 'E','L','G'
    std::vector<double> lcon, ucon, c, lvar, uvar;
    std::vector<char> marked;                          // integer-marked column
    std::vector<Coo> entries;
    bool have_obj = false;
    std::string rhs_name, rng_name, bnd_name;
    bool have_rhs = false, have_rng = false, have_bnd = false;
    double c0 = 0.0;
};

bool read_line(gzFile f, std::string &line) {
    line.clear();
    char buf[4096];
    bool any = false;
    while (gzgets(f, buf, sizeof(buf))) {
        any = true;
        line += buf;
        if (!line.empty() && line.back() == '\n') break;
    }
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    return any;
}

void split(const std::string &line, std::vector<std::string> &f) {
    f.clear();
    size_t i = 0, nline = line.size();
    while (i < nline && f.size() < 6) {
        while (i < nline && (line[i] == ' ' || line[i] == '\t')) ++i;
        if (i >= nline) break;
        size_t j = i;
        while (j < nline && line[j] != ' ' && line[j] != '\t') ++j;
        f.emplace_back(line, i, j - i);
        i = j;
    }
}

void set_coef(MpsData &d, int col, const std::string &rowname, double val, int lineno) {
    auto it = d.row_index.find(rowname);
    if (it == d.row_index.end()) {
        std::cerr << "Error: Unknown row " << rowname << " at line " << lineno << "\n";
        return;
    }
    if (it->second == 0) d.c[col] = val;
    else if (it->second > 0) d.entries.push_back({it->second - 1, col, val});
}

void set_rhs(MpsData &d, const std::string &rowname, double val, int lineno) {
    auto it = d.row_index.find(rowname);
    if (it == d.row_index.end()) { std::cerr << "Error: Unknown row " << rowname << "\n"; return; }
    const int row = it->second;
    if (row == 0) { d.c0 = -val; return; }
    if (row < 0) { std::cerr << "Error: Ignoring RHS for rim objective " << rowname << " at line " << lineno << "\n"; return; }
    const int k = row - 1;
    if (d.row_type[k] == 'E') { d.lcon[k] = val; d.ucon[k] = val; }
    else if (d.row_type[k] == 'L') d.ucon[k] = val;
    else if (d.row_type[k] == 'G') d.lcon[k] = val;
}

void set_range(MpsData &d, const std::string &rowname, double val, int lineno) {
    auto it = d.row_index.find(rowname);
    if (it == d.row_index.end()) { std::cerr << "Error: Unknown row " << rowname << " in RANGES section (l. " << lineno << ")\n"; return; }
    const int row = it->second;
    if (row <= 0) { std::cerr << "Error: Encountered objective row " << rowname << " in RANGES section (l. " << lineno << ")\n"; return; }
    const int k = row - 1;
    if (d.row_type[k] == 'E') { if (val >= 0.0) d.ucon[k] += val; else d.lcon[k] += val; }
    else if (d.row_type[k] == 'L') d.lcon[k] = d.ucon[k] - std::fabs(val);
    else if (d.row_type[k] == 'G') d.ucon[k] = d.lcon[k] + std::fabs(val);
}

}  // namespace

bool build_model_from_mps(const char *path, LP_info_cpu *lp) {
    std::printf("Start reading file....\n");
    const clock_t t0 = clock();
    gzFile f = gzopen(path, "rb");   // transparently reads plain and .gz files
    if (!f) {
        std::cerr << "Error: Cannot open file " << path << "\n";
        std::cerr << "Error: Failed to read MPS file\n";
        return false;
    }
    MpsData d;
    Section sec = S_NONE;
    bool integer_section = false, seen_rows = false, seen_cols = false, endata = false;
    std::string line;
    std::vector<std::string> fld;
    int lineno = 0;
    const double NaN = std::nan("");
    while (read_line(f, line)) {
        ++lineno;
        if (line.empty() || line[0] == '*' || line[0] == '&') continue;
        if (line[0] != ' ' && line[0] != '\t') {   // section header
            split(line, fld);
            if (fld.empty()) continue;
            const std::string &h = fld[0];
            if (h == "ENDATA") { endata = true; break; }
            else if (h == "NAME") { /* name ignored */ }
            else if (h == "OBJSENSE") sec = S_OBJSENSE;
            else if (h == "ROWS") { sec = S_ROWS; seen_rows = true; }
            else if (h == "COLUMNS") {
                if (!seen_rows) { std::cerr << "Error: ROWS section must come before COLUMNS\n"; gzclose(f); return false; }
                sec = S_COLUMNS; seen_cols = true;
            } else if (h == "RHS") {
                if (!seen_rows || !seen_cols) { std::cerr << "Error: RHS section must come after ROWS and COLUMNS\n"; gzclose(f); return false; }
                sec = S_RHS;
            } else if (h == "BOUNDS") {
                if (!seen_cols) { std::cerr << "Error: BOUNDS section must come after COLUMNS\n"; gzclose(f); return false; }
                sec = S_BOUNDS;
            } else if (h == "RANGES") {
                if (!seen_rows || !seen_cols) { std::cerr << "Error: RANGES section must come after ROWS and COLUMNS\n"; gzclose(f); return false; }
                sec = S_RANGES;
            } else if (h == "QUADOBJ" || h == "QMATRIX" || h == "OBJECT") sec = S_OTHER;
            // unknown headers keep the current section, as in the reference (SECTION_NONE falls through)
            continue;
        }
        split(line, fld);
        const int nf = (int)fld.size();
        if (nf == 0) continue;
        switch (sec) {
            case S_ROWS: {
                if (nf < 2) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                const std::string &t = fld[0], &name = fld[1];
                const bool is_con = (t == "E" || t == "L" || t == "G");
                if (!is_con) {   // N (or anything else) is an objective row
                    if (!d.have_obj) { d.have_obj = true; d.row_index[name] = 0; }
                    else { std::cerr << "Warning: Detected rim objective row " << name << " at line " << lineno << "\n"; d.row_index[name] = -1; }
                    break;
                }
                const int k = (int)d.row_type.size();
                d.row_index[name] = k + 1;
                d.row_type.push_back(t[0]);
                if (t == "E") { d.lcon.push_back(0.0); d.ucon.push_back(0.0); }
                else if (t == "G") { d.lcon.push_back(0.0); d.ucon.push_back(INFINITY); }
                else { d.lcon.push_back(-INFINITY); d.ucon.push_back(0.0); }
                break;
            }
            case S_COLUMNS: {
                if (nf >= 3 && fld[1] == "'MARKER'") {
                    if (fld[2] == "'INTORG'") integer_section = true;
                    else if (fld[2] == "'INTEND'") integer_section = false;
                    else std::cerr << "Error: Ignoring marker " << fld[2] << " at line " << lineno << "\n";
                    break;
                }
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                int col;
                auto it = d.col_index.find(fld[0]);
                if (it == d.col_index.end()) {
                    col = (int)d.c.size();
                    d.col_index.emplace(fld[0], col);
                    d.c.push_back(0.0); d.lvar.push_back(NaN); d.uvar.push_back(NaN);
                    d.marked.push_back(integer_section ? 1 : 0);
                } else col = it->second;
                set_coef(d, col, fld[1], atof(fld[2].c_str()), lineno);
                if (nf >= 5) set_coef(d, col, fld[3], atof(fld[4].c_str()), lineno);
                break;
            }
            case S_RHS: {
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                if (!d.have_rhs) { d.have_rhs = true; d.rhs_name = fld[0]; }
                else if (d.rhs_name != fld[0]) { std::cerr << "Error: Skipping line " << lineno << " with rim RHS " << fld[0] << "\n"; break; }
                set_rhs(d, fld[1], atof(fld[2].c_str()), lineno);
                if (nf >= 5) set_rhs(d, fld[3], atof(fld[4].c_str()), lineno);
                break;
            }
            case S_RANGES: {
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                if (!d.have_rng) { d.have_rng = true; d.rng_name = fld[0]; }
                else if (d.rng_name != fld[0]) { std::cerr << "Error: Skipping line " << lineno << " with rim RANGES " << fld[0] << "\n"; break; }
                set_range(d, fld[1], atof(fld[2].c_str()), lineno);
                if (nf >= 5 && !fld[3].empty()) set_range(d, fld[3], atof(fld[4].c_str()), lineno);
                break;
            }
            case S_BOUNDS: {
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                if (!d.have_bnd) { d.have_bnd = true; d.bnd_name = fld[1]; }
                else if (d.bnd_name != fld[1]) { std::cerr << "Error: Skipping line " << lineno << " with rim bound " << fld[1] << "\n"; break; }
                auto it = d.col_index.find(fld[2]);
                if (it == d.col_index.end()) { std::cerr << "Error: Unknown column " << fld[2] << "\n"; break; }
                const int col = it->second;
                const std::string &bt = fld[0];
                if (bt == "FR") { d.lvar[col] = -INFINITY; d.uvar[col] = INFINITY; break; }
                if (bt == "MI") { d.lvar[col] = -INFINITY; break; }
                if (bt == "PL") { d.uvar[col] = INFINITY; break; }
                if (bt == "BV") { d.lvar[col] = 0.0; d.uvar[col] = 1.0; break; }
                if (nf < 4) { std::cerr << "Error: At least 4 fields required for " << bt << " bounds\n"; break; }
                const double val = atof(fld[3].c_str());
                if (bt == "LO" || bt == "LI") d.lvar[col] = val;
                else if (bt == "UP" || bt == "UI") d.uvar[col] = val;
                else if (bt == "FX") { d.lvar[col] = val; d.uvar[col] = val; }
                else std::cerr << "Warning: Unknown bound type " << bt << "\n";
                break;
            }
            default: break;   // OBJSENSE parsed and ignored (never applied by the reference), QUADOBJ etc. skipped
        }
    }
    gzclose(f);
    if (!endata) std::cerr << "Warning: Reached end of file before ENDATA section\n";

    const int n = (int)d.c.size(), m = (int)d.row_type.size();
    for (int j = 0; j < n; ++j) {   // default bounds, reference src/mps_reader.cpp:1150-1180
        const bool ln = std::isnan(d.lvar[j]), un = std::isnan(d.uvar[j]);
        if (ln && un) { d.lvar[j] = 0.0; d.uvar[j] = d.marked[j] ? 1.0 : INFINITY; }
        else if (ln && !un) d.lvar[j] = (d.uvar[j] < 0) ? -INFINITY : 0.0;
        else if (!ln && un) d.uvar[j] = INFINITY;
    }
    std::printf("File reading time: %.4f seconds\n", (double)(clock() - t0) / CLOCKS_PER_SEC);

    // COO -> CSR: sort by (row, col), sum duplicates
    std::sort(d.entries.begin(), d.entries.end(), [](const Coo &a, const Coo &b) {
        return a.row != b.row ? a.row < b.row : a.col < b.col;
    });
    std::vector<int> cols; std::vector<double> vals;
    cols.reserve(d.entries.size()); vals.reserve(d.entries.size());
    for (size_t k = 0; k < d.entries.size(); ++k) {
        if (k > 0 && d.entries[k].row == d.entries[k - 1].row && d.entries[k].col == d.entries[k - 1].col) {
            vals.back() += d.entries[k].val;
        } else {
            cols.push_back(d.entries[k].col); vals.push_back(d.entries[k].val);
        }
    }
    // Row pointers: the reference derives them from the first unique_nnz entries of the SORTED,
    // NOT-YET-DEDUPLICATED list (src/mps_reader.cpp:1336-1355).  Without duplicate (row,col) cards this
    // is the ordinary CSR row pointer; with duplicates it mis-assigns row boundaries.  Mirrored exactly
    // so that the model arrays stay bit-identical to the reference on every input (DESIGN.md quirk list).
    std::vector<int> rowptr((size_t)m + 1, 0);
    {
        int row = 0;
        for (size_t i = 0; i < vals.size(); ++i) {
            const int entry_row = d.entries[i].row;
            while (row < entry_row) { row++; rowptr[row] = (int)i; }
        }
        while (row < m) { row++; rowptr[row] = (int)vals.size(); }
    }
    const int nnz = (int)vals.size();
    if (m <= 0 || n <= 0 || nnz <= 0) {
        std::cerr << "Error: Invalid dimensions in build_model_from_arrays: m=" << m << ", n=" << n << ", nnz=" << nnz << std::endl;
        return false;
    }
    std::printf("problem information: nRow = %d, nCol = %d, nnz A = %d\n\n", m, n, nnz);
    lp->m = m; lp->n = n; lp->obj_constant = d.c0;
    lp->A = static_cast<sparseMatrix *>(std::malloc(sizeof(sparseMatrix)));
    lp->A->row = m; lp->A->col = n; lp->A->numElements = nnz;
    lp->A->rowPtr = static_cast<int *>(std::malloc(sizeof(int) * ((size_t)m + 1)));
    lp->A->colIndex = static_cast<int *>(std::malloc(sizeof(int) * (size_t)nnz));
    lp->A->value = static_cast<double *>(std::malloc(sizeof(double) * (size_t)nnz));
    std::memcpy(lp->A->rowPtr, rowptr.data(), sizeof(int) * ((size_t)m + 1));
    std::memcpy(lp->A->colIndex, cols.data(), sizeof(int) * (size_t)nnz);
    std::memcpy(lp->A->value, vals.data(), sizeof(double) * (size_t)nnz);
    auto dup = [](const std::vector<double> &v) {
        double *p = static_cast<double *>(std::malloc(sizeof(double) * std::max<size_t>(v.size(), 1)));
        std::memcpy(p, v.data(), sizeof(double) * v.size());
        return p;
    };
    lp->AL = dup(d.lcon); lp->AU = dup(d.ucon); lp->c = dup(d.c); lp->l = dup(d.lvar); lp->u = dup(d.uvar);
    return true;
}

}  // namespace hpr
