// solve_mps_file -- command-line front end of libhprlp: build/solve_mps_file -i file.mps[.gz] [options]
//
// Same command line as the reference's CLI (src/solve_mps_file.cpp:14-134), so scripts written for it keep working:
// same flags, same rule for boolean values (only "true" and "1" switch a flag on, anything else switches it off),
// the same diagnostics and exit codes (0 after any solve, 1 for usage / missing value / missing file / parse
// failure), and no output besides what create_model_from_mps and solve print themselves.
// --cusparse-spmv and --autotune-verbose are accepted and have no effect (one hand-written backend).
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <string>

#include "../../include/HPRLP.h"
#include "../../include/hprlp_b200.h"

namespace {

void usage(const char *prog) {
    std::printf(
        "Usage: %s -i <input.mps|input.mps.gz> [options]\n"
        "Options:\n"
        "  -i, --input <path>         Path to input .mps or .mps.gz file (required)\n"
        "      --device <id>          CUDA device id (default: 0)\n"
        "      --max-iter <N>         Max iterations (default: INT32_MAX)\n"
        "      --tol <eps>            Stopping tolerance (default: 1e-4)\n"
        "      --time-limit <sec>     Time limit in seconds (default: 3600)\n"
        "      --check-iter <N>       Check interval (default: 150)\n"
        "      --cusparse-spmv <true/false>     accepted, no effect (single hand-written backend)\n"
        "      --autotune-verbose <true/false>  accepted, no effect\n"
        "      --cr <true/false>      Enable/disable Curtis-Reid prescaling (default: true)\n"
        "      --ruiz <true/false>    Enable/disable Ruiz scaling (default: true)\n"
        "      --pock <true/false>    Enable/disable Pock-Chambolle scaling (default: true)\n"
        "      --bc <true/false>      Enable/disable bounds/cost scaling (default: true)\n"
        "      --presolve <true/false>  Enable/disable embedded PSLP presolve (default: true)\n"
        "  -h, --help                 Show this help and exit\n"
        "\nExample:\n  %s -i model.mps.gz --device 0 --time-limit 3600 --tol 1e-4\n",
        prog, prog);
    std::fflush(stdout);
}

bool on(const std::string &v) { return v == "true" || v == "1"; }

}  // namespace

int main(int argc, char **argv) {
    HPRLP_parameters param;   // defaults of include/structs.h
    std::string input;
    bool have_input = false;
    // option -> action on its value
    const std::map<std::string, std::function<void(const std::string &)>> opts = {
        {"-i", [&](const std::string &v) { input = v; have_input = true; }},
        {"--input", [&](const std::string &v) { input = v; have_input = true; }},
        {"--device", [&](const std::string &v) { param.device_number = std::stoi(v); }},
        {"--max-iter", [&](const std::string &v) { param.max_iter = std::stoi(v); }},
        {"--tol", [&](const std::string &v) { param.stop_tol = std::stod(v); }},
        {"--time-limit", [&](const std::string &v) { param.time_limit = std::stod(v); }},
        {"--check-iter", [&](const std::string &v) { param.check_iter = std::stoi(v); }},
        {"--cusparse-spmv", [&](const std::string &v) { param.CUSPARSE_spmv = on(v); }},
        {"--autotune-verbose", [&](const std::string &v) { param.autotune_verbose = on(v); }},
        {"--cr", [&](const std::string &v) { param.use_CR_scaling = on(v); }},
        {"--ruiz", [&](const std::string &v) { param.use_Ruiz_scaling = on(v); }},
        {"--pock", [&](const std::string &v) { param.use_Pock_Chambolle_scaling = on(v); }},
        {"--bc", [&](const std::string &v) { param.use_bc_scaling = on(v); }},
        {"--presolve", [&](const std::string &v) { param.use_presolve = on(v); }},
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
        const auto it = opts.find(a);
        if (it == opts.end()) {
            std::fprintf(stderr, "Unknown option: %s\n", a.c_str());
            usage(argv[0]);
            return 1;
        }
        if (i + 1 >= argc) {
            std::fprintf(stderr, "Missing value for option: %s\n", a.c_str());
            usage(argv[0]);
            return 1;
        }
        it->second(argv[++i]);
    }
    if (!have_input) {
        std::fprintf(stderr, "Error: Input file is required. Use -i or --input option.\n");
        usage(argv[0]);
        return 1;
    }
    struct stat sb;
    if (stat(input.c_str(), &sb) != 0) {
        std::fprintf(stderr, "Input file does not exist: %s\n", input.c_str());
        usage(argv[0]);
        return 1;
    }
    hprlp_b200_warmup(param.device_number);   // CUDA context + cuRAND come up while the MPS file is parsed
    LP_info_cpu *model = create_model_from_mps(input.c_str());
    if (!model) {
        std::fprintf(stderr, "Failed to load model from MPS file: %s\n", input.c_str());
        return 1;
    }
    HPRLP_results out = solve(model, &param);
    std::free(out.x);
    std::free(out.y);
    std::free(out.z);
    free_model(model);
    return 0;
}
