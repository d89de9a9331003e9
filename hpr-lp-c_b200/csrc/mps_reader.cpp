// mps_reader.cpp -- host-side free-format MPS reader behind create_model_from_mps / build/solve_mps_file
// (for large files the reference's single-threaded fgets/atof/qsort reader takes longer than
// the solve).  Behaviour mirrors the reference reader (src/mps_reader.cpp) so that the resulting CSR / bound / cost
// arrays are bit-identical (tests/test_model_layer.py compares them byte for byte with the reference build):
//   * files are read as FREE format only (src/mps_reader.cpp:1517); '*' and '&' lines are comments;
//   * first N row is the objective, later N rows are ignored ("rim") (:598-612);
//   * RHS on the objective row sets obj_constant = -value (:765-767);
//   * RANGES: E rows take the sign of the range, L/G rows use |range| (:808-836);
//   * bounds default to [0,inf), [0,1] for integer-marked columns; an upper bound < 0 with no
//     lower bound gives l = -inf (:1150-1180); OBJSENSE is parsed and ignored (quirk 12);
//   * COO -> CSR sorts by (row, col) and sums duplicates (:1266-1361); numbers as atof reads them.
// Design (written from scratch): the file is inflated/read into one buffer; names are string views into it, kept
// in open-addressing hash tables (no per-field allocation); numbers go through std::from_chars (correctly rounded,
// same doubles as atof) with an strtod fallback for the spellings from_chars rejects; the COLUMNS section -- all
// of the volume -- is parsed by all host threads: a first parallel sweep finds the runs of equal column names, a
// short sequential pass numbers the columns in order of first appearance (and tracks the integer markers), a
// second parallel sweep converts the entries; COO -> CSR is a counting sort by row + per-row stable sorts, in
// parallel.  The small sections (ROWS, RHS, RANGES, BOUNDS) stay sequential.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "engine.h"

namespace hpr {
namespace {

enum Section { S_NONE, S_OBJSENSE, S_ROWS, S_COLUMNS, S_RHS, S_BOUNDS, S_RANGES, S_OTHER };

struct Tok {
    const char *p = nullptr;
    int len = 0;
    bool is(const char *s) const { return (int)std::strlen(s) == len && std::memcmp(p, s, len) == 0; }
    bool same(const Tok &o) const { return len == o.len && std::memcmp(p, o.p, len) == 0; }
    std::string str() const { return std::string(p, (size_t)len); }
};

// name -> int, keys are views into the file buffer
class NameMap {
  public:
    NameMap() { rehash(1024); }
    int *find(const Tok &k) {
        size_t i = hash(k) & mask_;
        while (slots_[i].p) {
            if (slots_[i].len == k.len && std::memcmp(slots_[i].p, k.p, k.len) == 0) return &slots_[i].val;
            i = (i + 1) & mask_;
        }
        return nullptr;
    }
    const int *find(const Tok &k) const { return const_cast<NameMap *>(this)->find(k); }
    void assign(const Tok &k, int v) {   // insert or overwrite (a repeated ROWS name re-binds, as operator[] does)
        if (int *p = find(k)) { *p = v; return; }
        if ((count_ + 1) * 2 > slots_.size()) rehash(slots_.size() * 2);
        place(k, hash(k), v);
        ++count_;
    }
    // The same with the hash computed elsewhere (the COLUMNS sweep hashes the names in parallel) and the table sized up
    // front, so that the one sequential pass over ~n column names neither rehashes nor waits on a cold slot.
    static size_t hash(const Tok &k) {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < k.len; ++i) { h ^= (unsigned char)k.p[i]; h *= 1099511628211ull; }
        return (size_t)(h ^ (h >> 29));
    }
    void reserve(size_t names) {
        size_t cap = slots_.size();
        while ((count_ + names + 1) * 2 > cap) cap *= 2;
        if (cap != slots_.size()) rehash(cap);
    }
    void prefetch(size_t h) const { __builtin_prefetch(&slots_[h & mask_]); }
    int *find(const Tok &k, size_t h) {
        size_t i = h & mask_;
        while (slots_[i].p) {
            if (slots_[i].len == k.len && std::memcmp(slots_[i].p, k.p, k.len) == 0) return &slots_[i].val;
            i = (i + 1) & mask_;
        }
        return nullptr;
    }
    void insert_new(const Tok &k, size_t h, int v) {   // k is known to be absent
        if ((count_ + 1) * 2 > slots_.size()) rehash(slots_.size() * 2);
        place(k, h, v);
        ++count_;
    }

  private:
    struct Slot { const char *p = nullptr; int len = 0; int val = 0; };
    void place(const Tok &k, size_t h, int v) {
        size_t i = h & mask_;
        while (slots_[i].p) i = (i + 1) & mask_;
        slots_[i].p = k.p; slots_[i].len = k.len; slots_[i].val = v;
    }
    void rehash(size_t cap) {
        std::vector<Slot> old;
        old.swap(slots_);
        slots_.assign(cap, Slot());
        mask_ = cap - 1;
        for (const Slot &s : old)
            if (s.p) place(Tok{s.p, s.len}, hash(Tok{s.p, s.len}), s.val);
    }
    std::vector<Slot> slots_;
    size_t mask_ = 0, count_ = 0;
};

struct Coo {
    int row, col;
    double val;
    Coo() {}   // deliberately uninitialised: vectors of entries are resized and then filled by several threads (first touch in parallel)
    Coo(int r, int c, double v) : row(r), col(c), val(v) {}
};

struct MpsData {
    NameMap row_index;                                 // 0 objective, -1 rim objective, k+1 constraint k
    NameMap col_index;
    std::vector<char> row_type;                        // 'E','L','G'
    std::vector<double> lcon, ucon, c, lvar, uvar;
    std::vector<char> marked;                          // integer-marked column
    std::vector<Coo> entries;
    bool have_obj = false;
    Tok rhs_name, rng_name, bnd_name;
    bool have_rhs = false, have_rng = false, have_bnd = false;
    double c0 = 0.0;
};

// up to 6 blank-separated fields of the line [b, e)
inline int split(const char *b, const char *e, Tok *f) {
    int n = 0;
    const char *i = b;
    while (i < e && n < 6) {
        while (i < e && (*i == ' ' || *i == '\t')) ++i;
        if (i >= e) break;
        const char *j = i;
        while (j < e && *j != ' ' && *j != '\t') ++j;
        f[n].p = i; f[n].len = (int)(j - i);
        ++n;
        i = j;
    }
    return n;
}

// atof(token): from_chars covers plain decimal / exponent spellings (and is correctly rounded like strtod);
// anything it does not consume completely -- leading '+', hex, "1.0D3" prefixes, junk -- goes through strtod on a
// terminated copy, which is what atof does.
inline double to_double(const Tok &t) {
    double v = 0.0;
    const char *b = t.p, *e = t.p + t.len;
    const auto r = std::from_chars(b, e, v, std::chars_format::general);
    if (r.ec == std::errc() && r.ptr == e) return v;
    char tmp[128];
    if (t.len < (int)sizeof(tmp)) {
        std::memcpy(tmp, t.p, t.len);
        tmp[t.len] = 0;
        return atof(tmp);
    }
    return atof(t.str().c_str());
}

// line [b, e) without the terminator; returns the start of the next line
inline const char *line_end(const char *b, const char *end, const char **e) {
    const char *nl = static_cast<const char *>(std::memchr(b, '\n', (size_t)(end - b)));
    const char *stop = nl ? nl : end;
    const char *next = nl ? nl + 1 : end;
    while (stop > b && (stop[-1] == '\r' || stop[-1] == '\n')) --stop;
    *e = stop;
    return next;
}
inline bool is_skip(const char *b, const char *e) { return b == e || *b == '*' || *b == '&'; }
inline bool is_header(const char *b, const char *e) { return !is_skip(b, e) && *b != ' ' && *b != '\t'; }

void set_rhs(MpsData &d, const Tok &rowname, double val, int lineno) {
    const int *it = d.row_index.find(rowname);
    if (!it) { std::cerr << "Error: Unknown row " << rowname.str() << "\n"; return; }
    const int row = *it;
    if (row == 0) { d.c0 = -val; return; }
    if (row < 0) { std::cerr << "Error: Ignoring RHS for rim objective " << rowname.str() << " at line " << lineno << "\n"; return; }
    const int k = row - 1;
    if (d.row_type[k] == 'E') { d.lcon[k] = val; d.ucon[k] = val; }
    else if (d.row_type[k] == 'L') d.ucon[k] = val;
    else if (d.row_type[k] == 'G') d.lcon[k] = val;
}

void set_range(MpsData &d, const Tok &rowname, double val, int lineno) {
    const int *it = d.row_index.find(rowname);
    if (!it) { std::cerr << "Error: Unknown row " << rowname.str() << " in RANGES section (l. " << lineno << ")\n"; return; }
    const int row = *it;
    if (row <= 0) { std::cerr << "Error: Encountered objective row " << rowname.str() << " in RANGES section (l. " << lineno << ")\n"; return; }
    const int k = row - 1;
    if (d.row_type[k] == 'E') { if (val >= 0.0) d.ucon[k] += val; else d.lcon[k] += val; }
    else if (d.row_type[k] == 'L') d.lcon[k] = d.ucon[k] - std::fabs(val);
    else if (d.row_type[k] == 'G') d.ucon[k] = d.lcon[k] + std::fabs(val);
}

int host_threads() {
#ifdef _OPENMP
    return std::max(1, omp_get_max_threads());
#else
    return 1;
#endif
}

// ---- COLUMNS block [b, e): whole lines, no section header inside -------------------------------------------------
struct ColEvent {
    enum Kind { RUN, INTORG, INTEND, BAD_MARKER, SHORT } kind;
    Tok name;          // RUN: column name; BAD_MARKER: the marker text
    const char *line;  // start of the (first) line
    int lineno;        // line number inside the block (1-based), made global by the caller
    int nf;            // SHORT: number of fields
    int col;           // RUN: column index (sequential pass)
    size_t hash;       // RUN: NameMap::hash(name), computed in the parallel sweep
};
struct ColChunk {
    const char *b, *e;
    int lines = 0;
    std::vector<ColEvent> events;
    std::vector<Coo> entries;
    std::vector<std::pair<int, double>> obj;                 // (column, objective coefficient) in line order
    std::vector<std::pair<int, std::string>> unknown_rows;   // (block line number, name)
};

double wall() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
const bool g_timing = getenv("HPRLP_TIMING") != nullptr;

void parse_columns_block(MpsData &d, const char *b, const char *e, int lineno_base, int *lines_out, bool *integer_section) {
    const double tt0 = wall();
    const int T = (int)std::max<long long>(1, std::min<long long>(host_threads(), (e - b) / (1 << 16) + 1));
    std::vector<ColChunk> ch(T);
    for (int t = 0; t < T; ++t) {   // chunk starts moved to the next line start
        const char *s = b + (e - b) * (long long)t / T;
        if (t > 0) {
            const char *nl = static_cast<const char *>(std::memchr(s - 1, '\n', (size_t)(e - (s - 1))));
            s = nl ? nl + 1 : e;
        }
        ch[t].b = s;
        if (t > 0) ch[t - 1].e = s;
    }
    ch[T - 1].e = e;

    // sweep 1 (parallel): runs of equal column names, markers, short lines; line counts
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        ColChunk &c = ch[t];
        Tok f[6], cur;
        bool have_cur = false;
        const char *p = c.b;
        int ln = 0;
        while (p < c.e) {
            const char *le;
            const char *next = line_end(p, c.e, &le);
            ++ln;
            if (!is_skip(p, le)) {
                const int nf = split(p, le, f);
                if (nf > 0) {
                    if (nf >= 3 && f[1].is("'MARKER'")) {
                        ColEvent ev{f[2].is("'INTORG'") ? ColEvent::INTORG : (f[2].is("'INTEND'") ? ColEvent::INTEND : ColEvent::BAD_MARKER),
                                    f[2], p, ln, nf, -1, 0};
                        c.events.push_back(ev);
                    } else if (nf < 3) {
                        c.events.push_back(ColEvent{ColEvent::SHORT, Tok(), p, ln, nf, -1, 0});
                    } else if (!have_cur || !cur.same(f[0])) {
                        cur = f[0]; have_cur = true;
                        c.events.push_back(ColEvent{ColEvent::RUN, f[0], p, ln, nf, -1, NameMap::hash(f[0])});
                    }
                }
            }
            p = next;
        }
        c.lines = ln;
    }

    const double tt1 = wall();
    // sequential pass: column numbers in order of first appearance, integer markers (reference :977-1040)
    const double NaN = std::nan("");
    std::vector<int> base(T + 1, 0);
    for (int t = 0; t < T; ++t) base[t + 1] = base[t] + ch[t].lines;
    {   // at most one new column per event: size the table and the column arrays once
        size_t nev = 0;
        for (int t = 0; t < T; ++t) nev += ch[t].events.size();
        d.col_index.reserve(nev);
        d.c.reserve(d.c.size() + nev); d.lvar.reserve(d.lvar.size() + nev); d.uvar.reserve(d.uvar.size() + nev);
        d.marked.reserve(d.marked.size() + nev);
    }
    constexpr size_t kAhead = 16;   // slots of the events this far ahead are prefetched (the table is tens of MB)
    for (int t = 0; t < T; ++t) {
        std::vector<ColEvent> &evs = ch[t].events;
        for (size_t i = 0; i < evs.size(); ++i) {
            if (i + kAhead < evs.size() && evs[i + kAhead].kind == ColEvent::RUN) d.col_index.prefetch(evs[i + kAhead].hash);
            ColEvent &ev = evs[i];
            const int lineno = lineno_base + base[t] + ev.lineno;
            switch (ev.kind) {
                case ColEvent::INTORG: *integer_section = true; break;
                case ColEvent::INTEND: *integer_section = false; break;
                case ColEvent::BAD_MARKER: std::cerr << "Error: Ignoring marker " << ev.name.str() << " at line " << lineno << "\n"; break;
                case ColEvent::SHORT: std::cerr << "Error: Line " << lineno << " contains only " << ev.nf << " fields\n"; break;
                case ColEvent::RUN: {
                    if (int *it = d.col_index.find(ev.name, ev.hash)) ev.col = *it;
                    else {
                        ev.col = (int)d.c.size();
                        d.col_index.insert_new(ev.name, ev.hash, ev.col);
                        d.c.push_back(0.0); d.lvar.push_back(NaN); d.uvar.push_back(NaN);
                        d.marked.push_back(*integer_section ? 1 : 0);
                    }
                    break;
                }
            }
        }
    }

    // sweep 2 (parallel): entries
    const double tt2 = wall();
    const MpsData &cd = d;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        ColChunk &c = ch[t];
        Tok f[6];
        size_t ev = 0;
        int col = -1;
        // a chunk may start inside a run begun in an earlier chunk: its first data line then opens a RUN event of its own
        // (have_cur was false), so col is always set before the first entry
        const char *p = c.b;
        int ln = 0;
        c.entries.reserve((size_t)(c.e - c.b) / 24);
        while (p < c.e) {
            const char *le;
            const char *next = line_end(p, c.e, &le);
            ++ln;
            bool data = true;
            while (ev < c.events.size() && c.events[ev].line == p) {
                if (c.events[ev].kind == ColEvent::RUN) col = c.events[ev].col;
                else data = false;   // marker or short line: nothing to convert
                ++ev;
            }
            if (data && !is_skip(p, le)) {
                const int nf = split(p, le, f);
                if (nf >= 3) {
                    for (int k = 1; k + 1 < nf && k <= 3; k += 2) {
                        if (k == 3 && nf < 5) break;
                        const int *it = cd.row_index.find(f[k]);
                        if (!it) { c.unknown_rows.emplace_back(ln, f[k].str()); continue; }
                        const double v = to_double(f[k + 1]);
                        if (*it == 0) c.obj.emplace_back(col, v);
                        else if (*it > 0) c.entries.push_back(Coo{*it - 1, col, v});
                    }
                }
            }
            p = next;
        }
    }
    const double tt3 = wall();
    std::vector<size_t> off(T + 1, d.entries.size());
    for (int t = 0; t < T; ++t) off[t + 1] = off[t] + ch[t].entries.size();
    d.entries.resize(off[T]);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t)
        if (!ch[t].entries.empty()) std::memcpy(static_cast<void *>(d.entries.data() + off[t]), ch[t].entries.data(), sizeof(Coo) * ch[t].entries.size());
    for (int t = 0; t < T; ++t) {
        for (const auto &o : ch[t].obj) d.c[o.first] = o.second;
        for (const auto &u : ch[t].unknown_rows)
            std::cerr << "Error: Unknown row " << u.second << " at line " << (lineno_base + base[t] + u.first) << "\n";
        std::vector<Coo>().swap(ch[t].entries);
    }
    *lines_out = base[T];
    if (g_timing)
        std::fprintf(stderr, "[hprlp timing] COLUMNS block %.1f MB, %d threads: runs %.3f s, numbering %.3f s, entries %.3f s, merge %.3f s\n",
                     (e - b) / 1e6, T, tt1 - tt0, tt2 - tt1, tt3 - tt2, wall() - tt3);
}

// first section header at or after p (parallel scan); `end` if there is none
const char *next_header(const char *p, const char *end) {
    const int T = (int)std::max<long long>(1, std::min<long long>(host_threads(), (end - p) / (1 << 20) + 1));
    std::vector<const char *> found(T, nullptr);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        const char *s = p + (end - p) * (long long)t / T;
        const char *stop = p + (end - p) * (long long)(t + 1) / T;
        if (t > 0) {
            const char *nl = static_cast<const char *>(std::memchr(s - 1, '\n', (size_t)(end - (s - 1))));
            s = nl ? nl + 1 : end;
        }
        while (s < stop) {   // lines STARTING in [s, stop)
            const char *le;
            const char *next = line_end(s, end, &le);
            if (is_header(s, le)) { found[t] = s; break; }
            s = next;
        }
    }
    for (int t = 0; t < T; ++t)
        if (found[t]) return found[t];
    return end;
}

// The whole file as one read-only byte range: plain files are mapped (no copy), gzip files (magic 1f 8b) are inflated
// through zlib into a heap buffer.
struct FileBytes {
    const char *data = nullptr;
    size_t size = 0;
    void *map = nullptr;
    size_t map_len = 0;
    std::vector<char> heap;
    ~FileBytes() { if (map) munmap(map, map_len); }
};

bool read_whole(const char *path, FileBytes &fb) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    unsigned char magic[2] = {0, 0};
    const ssize_t got2 = pread(fd, magic, 2, 0);
    const bool gz = got2 == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    struct stat st;
    if (!gz && fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
        if (st.st_size == 0) { close(fd); return true; }
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (m != MAP_FAILED) {
            close(fd);
            madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
            fb.map = m; fb.map_len = (size_t)st.st_size;
            fb.data = static_cast<const char *>(m); fb.size = (size_t)st.st_size;
            return true;
        }
    }
    close(fd);
    gzFile f = gzopen(path, "rb");   // gzip (or a non-regular plain stream: zlib reads those transparently)
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    size_t cap = (size_t)1 << 24, len = 0;
    fb.heap.resize(cap);
    for (;;) {
        if (len == cap) { cap *= 2; fb.heap.resize(cap); }
        const int want = (int)std::min<size_t>(cap - len, (size_t)1 << 30);
        const int got = gzread(f, fb.heap.data() + len, (unsigned)want);
        if (got <= 0) break;
        len += (size_t)got;
    }
    gzclose(f);
    fb.heap.resize(len);
    fb.data = fb.heap.data(); fb.size = len;
    return true;
}

}  // namespace

bool build_model_from_mps(const char *path, LP_info_cpu *lp) {
    std::printf("Start reading file....\n");
    const auto t0 = std::chrono::steady_clock::now();   // wall clock (the parse is multi-threaded; clock() would add the threads up)
    FileBytes buf;
    if (!read_whole(path, buf)) {
        std::cerr << "Error: Cannot open file " << path << "\n";
        std::cerr << "Error: Failed to read MPS file\n";
        return false;
    }
    if (g_timing) std::fprintf(stderr, "[hprlp timing] file read %.3f s (%.1f MB)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), buf.size / 1e6);
    const char *p = buf.data, *end = buf.data + buf.size;
    MpsData d;
    Section sec = S_NONE;
    bool integer_section = false, seen_rows = false, seen_cols = false, endata = false;
    Tok fld[6];
    int lineno = 0;
    while (p < end) {
        const char *le;
        const char *next = line_end(p, end, &le);
        if (is_skip(p, le)) { ++lineno; p = next; continue; }
        if (is_header(p, le)) {
            ++lineno;
            const int nf = split(p, le, fld);
            p = next;
            if (nf == 0) continue;
            const Tok &h = fld[0];
            if (h.is("ENDATA")) { endata = true; break; }
            else if (h.is("NAME")) { /* name ignored */ }
            else if (h.is("OBJSENSE")) sec = S_OBJSENSE;
            else if (h.is("ROWS")) { sec = S_ROWS; seen_rows = true; }
            else if (h.is("COLUMNS")) {
                if (!seen_rows) { std::cerr << "Error: ROWS section must come before COLUMNS\n"; return false; }
                sec = S_COLUMNS; seen_cols = true;
            } else if (h.is("RHS")) {
                if (!seen_rows || !seen_cols) { std::cerr << "Error: RHS section must come after ROWS and COLUMNS\n"; return false; }
                sec = S_RHS;
            } else if (h.is("BOUNDS")) {
                if (!seen_cols) { std::cerr << "Error: BOUNDS section must come after COLUMNS\n"; return false; }
                sec = S_BOUNDS;
            } else if (h.is("RANGES")) {
                if (!seen_rows || !seen_cols) { std::cerr << "Error: RANGES section must come after ROWS and COLUMNS\n"; return false; }
                sec = S_RANGES;
            } else if (h.is("QUADOBJ") || h.is("QMATRIX") || h.is("OBJECT")) sec = S_OTHER;
            // unknown headers keep the current section, as in the reference (SECTION_NONE falls through)
            continue;
        }
        if (sec == S_COLUMNS) {   // everything up to the next header, by all threads
            const char *blk_end = next_header(p, end);
            int lines = 0;
            parse_columns_block(d, p, blk_end, lineno, &lines, &integer_section);
            lineno += lines;
            p = blk_end;
            continue;
        }
        ++lineno;
        const int nf = split(p, le, fld);
        p = next;
        if (nf == 0) continue;
        switch (sec) {
            case S_ROWS: {
                if (nf < 2) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                const Tok &t = fld[0], &name = fld[1];
                const bool is_con = (t.is("E") || t.is("L") || t.is("G"));
                if (!is_con) {   // N (or anything else) is an objective row
                    if (!d.have_obj) { d.have_obj = true; d.row_index.assign(name, 0); }
                    else { std::cerr << "Warning: Detected rim objective row " << name.str() << " at line " << lineno << "\n"; d.row_index.assign(name, -1); }
                    break;
                }
                const int k = (int)d.row_type.size();
                d.row_index.assign(name, k + 1);
                d.row_type.push_back(t.p[0]);
                if (t.is("E")) { d.lcon.push_back(0.0); d.ucon.push_back(0.0); }
                else if (t.is("G")) { d.lcon.push_back(0.0); d.ucon.push_back(INFINITY); }
                else { d.lcon.push_back(-INFINITY); d.ucon.push_back(0.0); }
                break;
            }
            case S_RHS: {
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                if (!d.have_rhs) { d.have_rhs = true; d.rhs_name = fld[0]; }
                else if (!d.rhs_name.same(fld[0])) { std::cerr << "Error: Skipping line " << lineno << " with rim RHS " << fld[0].str() << "\n"; break; }
                set_rhs(d, fld[1], to_double(fld[2]), lineno);
                if (nf >= 5) set_rhs(d, fld[3], to_double(fld[4]), lineno);
                break;
            }
            case S_RANGES: {
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                if (!d.have_rng) { d.have_rng = true; d.rng_name = fld[0]; }
                else if (!d.rng_name.same(fld[0])) { std::cerr << "Error: Skipping line " << lineno << " with rim RANGES " << fld[0].str() << "\n"; break; }
                set_range(d, fld[1], to_double(fld[2]), lineno);
                if (nf >= 5 && fld[3].len > 0) set_range(d, fld[3], to_double(fld[4]), lineno);
                break;
            }
            case S_BOUNDS: {
                if (nf < 3) { std::cerr << "Error: Line " << lineno << " contains only " << nf << " fields\n"; break; }
                if (!d.have_bnd) { d.have_bnd = true; d.bnd_name = fld[1]; }
                else if (!d.bnd_name.same(fld[1])) { std::cerr << "Error: Skipping line " << lineno << " with rim bound " << fld[1].str() << "\n"; break; }
                const int *it = d.col_index.find(fld[2]);
                if (!it) { std::cerr << "Error: Unknown column " << fld[2].str() << "\n"; break; }
                const int col = *it;
                const Tok &bt = fld[0];
                if (bt.is("FR")) { d.lvar[col] = -INFINITY; d.uvar[col] = INFINITY; break; }
                if (bt.is("MI")) { d.lvar[col] = -INFINITY; break; }
                if (bt.is("PL")) { d.uvar[col] = INFINITY; break; }
                if (bt.is("BV")) { d.lvar[col] = 0.0; d.uvar[col] = 1.0; break; }
                if (nf < 4) { std::cerr << "Error: At least 4 fields required for " << bt.str() << " bounds\n"; break; }
                const double val = to_double(fld[3]);
                if (bt.is("LO") || bt.is("LI")) d.lvar[col] = val;
                else if (bt.is("UP") || bt.is("UI")) d.uvar[col] = val;
                else if (bt.is("FX")) { d.lvar[col] = val; d.uvar[col] = val; }
                else std::cerr << "Warning: Unknown bound type " << bt.str() << "\n";
                break;
            }
            default: break;   // OBJSENSE parsed and ignored (never applied by the reference), QUADOBJ etc. skipped
        }
    }
    if (!endata) std::cerr << "Warning: Reached end of file before ENDATA section\n";

    const int n = (int)d.c.size(), m = (int)d.row_type.size();
    for (int j = 0; j < n; ++j) {   // default bounds, reference src/mps_reader.cpp:1150-1180
        const bool ln = std::isnan(d.lvar[j]), un = std::isnan(d.uvar[j]);
        if (ln && un) { d.lvar[j] = 0.0; d.uvar[j] = d.marked[j] ? 1.0 : INFINITY; }
        else if (ln && !un) d.lvar[j] = (d.uvar[j] < 0) ? -INFINITY : 0.0;
        else if (!ln && un) d.uvar[j] = INFINITY;
    }
    std::printf("File reading time: %.4f seconds\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

    // COO -> CSR: sort by (row, col), sum duplicates.  Stable counting sort by row, by all threads (thread t scatters the
    // t-th slice of the file-ordered entries behind the slices before it), then every row stable-sorted by column when it
    // is not already in order; rows in parallel.
    const size_t ne = d.entries.size();
    std::vector<Coo> sorted(ne);
    std::vector<size_t> start((size_t)m + 1, 0);
    size_t n_dup = 0;
    {
        const int T = (int)std::max<long long>(1, std::min<long long>(host_threads(), (long long)(ne >> 16) + 1));
        std::vector<std::vector<unsigned>> cnt(T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; ++t) {
            cnt[t].assign((size_t)m, 0u);
            const size_t lo = ne * (size_t)t / T, hi = ne * (size_t)(t + 1) / T;
            for (size_t k = lo; k < hi; ++k) cnt[t][d.entries[k].row]++;
        }
        for (int i = 0; i < m; ++i) {   // start[i]: first slot of row i; cnt[t][i] becomes thread t's first slot in row i
            size_t run = start[i];
            for (int t = 0; t < T; ++t) { const unsigned c = cnt[t][i]; cnt[t][i] = (unsigned)(run - start[i]); run += c; }
            start[i + 1] = run;
        }
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; ++t) {
            const size_t lo = ne * (size_t)t / T, hi = ne * (size_t)(t + 1) / T;
            for (size_t k = lo; k < hi; ++k) {
                const int r = d.entries[k].row;
                sorted[start[r] + cnt[t][r]++] = d.entries[k];
            }
        }
        std::vector<Coo>().swap(d.entries);
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_dup)
        for (int i = 0; i < m; ++i) {
            Coo *rb = sorted.data() + start[i], *re = sorted.data() + start[i + 1];
            bool in_order = true;
            for (Coo *q = rb + 1; q < re; ++q)
                if (q->col < q[-1].col) { in_order = false; break; }
            if (!in_order) std::stable_sort(rb, re, [](const Coo &a, const Coo &b) { return a.col < b.col; });
            for (Coo *q = rb + 1; q < re; ++q)
                if (q->col == q[-1].col) ++n_dup;
        }
    }
    if (g_timing) std::fprintf(stderr, "[hprlp timing] sorted at %.3f s (%zu duplicate cards)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), n_dup);
    std::vector<int> cols; std::vector<double> vals;
    std::vector<int> rowptr((size_t)m + 1, 0);
    int *direct_cols = nullptr;      // no duplicate cards: the model's arrays are filled straight from the sorted list by all
    double *direct_vals = nullptr;   // threads (no intermediate vectors, no single-threaded 12 B/nnz copy)
    if (n_dup == 0) {
        // no duplicate (row, col) card: the sorted list IS the CSR matrix
        if (m > 0 && n > 0 && ne > 0) {
            direct_cols = static_cast<int *>(std::malloc(sizeof(int) * ne));
            direct_vals = static_cast<double *>(std::malloc(sizeof(double) * ne));
            if (!direct_cols || !direct_vals) { std::free(direct_cols); std::free(direct_vals); throw std::bad_alloc(); }
#pragma omp parallel for schedule(static)
            for (long long k = 0; k < (long long)ne; ++k) { direct_cols[k] = sorted[k].col; direct_vals[k] = sorted[k].val; }
        }
        for (int i = 0; i <= m; ++i) rowptr[i] = (int)start[i];
    } else {
        cols.reserve(ne); vals.reserve(ne);
        for (size_t k = 0; k < ne; ++k) {
            if (k > 0 && sorted[k].row == sorted[k - 1].row && sorted[k].col == sorted[k - 1].col) {
                vals.back() += sorted[k].val;
            } else {
                cols.push_back(sorted[k].col); vals.push_back(sorted[k].val);
            }
        }
        // Row pointers: the reference derives them from the first unique_nnz entries of the SORTED,
        // NOT-YET-DEDUPLICATED list (src/mps_reader.cpp:1336-1355).  Without duplicate (row,col) cards this
        // is the ordinary CSR row pointer (the branch above); with duplicates it mis-assigns row boundaries.  Mirrored
        // exactly so that the model arrays stay bit-identical to the reference on every input (DESIGN.md quirk list).
        int row = 0;
        for (size_t i = 0; i < vals.size(); ++i) {
            const int entry_row = sorted[i].row;
            while (row < entry_row) { row++; rowptr[row] = (int)i; }
        }
        while (row < m) { row++; rowptr[row] = (int)vals.size(); }
    }
    const int nnz = direct_cols ? (int)ne : (int)vals.size();
    if (m <= 0 || n <= 0 || nnz <= 0) {
        std::cerr << "Error: Invalid dimensions in build_model_from_arrays: m=" << m << ", n=" << n << ", nnz=" << nnz << std::endl;
        return false;
    }
    std::printf("problem information: nRow = %d, nCol = %d, nnz A = %d\n\n", m, n, nnz);
    lp->m = m; lp->n = n; lp->obj_constant = d.c0;
    lp->A = static_cast<sparseMatrix *>(std::malloc(sizeof(sparseMatrix)));
    lp->A->row = m; lp->A->col = n; lp->A->numElements = nnz;
    lp->A->rowPtr = static_cast<int *>(std::malloc(sizeof(int) * ((size_t)m + 1)));
    std::memcpy(lp->A->rowPtr, rowptr.data(), sizeof(int) * ((size_t)m + 1));
    if (direct_cols) {
        lp->A->colIndex = direct_cols;
        lp->A->value = direct_vals;
    } else {
        lp->A->colIndex = static_cast<int *>(std::malloc(sizeof(int) * (size_t)nnz));
        lp->A->value = static_cast<double *>(std::malloc(sizeof(double) * (size_t)nnz));
        std::memcpy(lp->A->colIndex, cols.data(), sizeof(int) * (size_t)nnz);
        std::memcpy(lp->A->value, vals.data(), sizeof(double) * (size_t)nnz);
    }
    auto dup = [](const std::vector<double> &v) {
        double *q = static_cast<double *>(std::malloc(sizeof(double) * std::max<size_t>(v.size(), 1)));
        std::memcpy(q, v.data(), sizeof(double) * v.size());
        return q;
    };
    lp->AL = dup(d.lcon); lp->AU = dup(d.ucon); lp->c = dup(d.c); lp->l = dup(d.lvar); lp->u = dup(d.uvar);
    return true;
}

}  // namespace hpr
