// Times create_model_from_mps of whichever libhprlp it is linked against (ours: lib/libhprlp.so; the reference build:
// oracle/_ref/libhprlp_ref.so).  g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include tools/mps_time.cpp -Llib -lhprlp
#include <chrono>
#include <cstdio>
#include "HPRLP.h"
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const auto t0 = std::chrono::steady_clock::now();
    LP_info_cpu *m = create_model_from_mps(argv[1]);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!m) return 1;
    std::fprintf(stderr, "create_model_from_mps wall %.3f s (m=%d n=%d nnz=%d)\n", s, m->m, m->n, m->A->numElements);
    free_model(m);
    return 0;
}
