// solve_mps_file -- command-line front end with the reference's flags
// (reference src/solve_mps_file.cpp:14-32): build/solve_mps_file -i file.mps [options]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/HPRLP.h"

static void usage(const char *prog) {
    std::printf("Usage: %s -i <mps_file> [options]\n", prog);
    std::printf("  -i, --input <file>        input MPS file (.mps or .mps.gz)\n");
    std::printf("  --device <id>             CUDA device (default 0)\n");
    std::printf("  --max-iter <n>            maximum iterations\n");
    std::printf("  --tol <eps>               stopping tolerance (default 1e-4)\n");
    std::printf("  --time-limit <sec>        time limit in seconds (default 3600)\n");
    std::printf("  --check-iter <n>          restart/check interval (default 150)\n");
    std::printf("  --cusparse-spmv <bool>    accepted for compatibility (single hand-written backend)\n");
    std::printf("  --autotune-verbose <bool> accepted for compatibility\n");
    std::printf("  --cr|--ruiz|--pock|--bc <bool>   scaling switches (default true)\n");
    std::printf("  --presolve <bool>         PSLP presolve (default true)\n");
}

static bool parse_bool(const char *s) {
    return !(std::strcmp(s, "0") == 0 || std::strcmp(s, "false") == 0 || std::strcmp(s, "False") == 0 ||
             std::strcmp(s, "off") == 0 || std::strcmp(s, "no") == 0);
}

int main(int argc, char **argv) {
    HPRLP_parameters p;
    std::string input;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return (i + 1 < argc) ? argv[++i] : ""; };
        if (a == "-i" || a == "--input") input = next();
        else if (a == "--device") p.device_number = std::atoi(next());
        else if (a == "--max-iter") p.max_iter = std::atoi(next());
        else if (a == "--tol") p.stop_tol = std::atof(next());
        else if (a == "--time-limit") p.time_limit = std::atof(next());
        else if (a == "--check-iter") p.check_iter = std::atoi(next());
        else if (a == "--cusparse-spmv") p.CUSPARSE_spmv = parse_bool(next());
        else if (a == "--autotune-verbose") p.autotune_verbose = parse_bool(next());
        else if (a == "--cr") p.use_CR_scaling = parse_bool(next());
        else if (a == "--ruiz") p.use_Ruiz_scaling = parse_bool(next());
        else if (a == "--pock") p.use_Pock_Chambolle_scaling = parse_bool(next());
        else if (a == "--bc") p.use_bc_scaling = parse_bool(next());
        else if (a == "--presolve") p.use_presolve = parse_bool(next());
        else if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
        else { std::fprintf(stderr, "Unknown option: %s\n", a.c_str()); usage(argv[0]); return 1; }
    }
    if (input.empty()) { usage(argv[0]); return 1; }
    LP_info_cpu *model = create_model_from_mps(input.c_str());
    if (!model) { std::fprintf(stderr, "Failed to create model from %s\n", input.c_str()); return 1; }
    HPRLP_results r = solve(model, &p);
    std::printf("Status: %s  iter: %d  primal_obj: %.10e  residual: %.3e  time: %.3f s\n", r.status, r.iter, r.primal_obj,
                r.residuals, r.time);
    std::free(r.x); std::free(r.y); std::free(r.z);
    free_model(model);
    return std::strcmp(r.status, "ERROR") == 0 ? 2 : 0;
}
