// Command line of the tiled_mm_b200 apps: the flags of the reference's miniapps (examples/multiply.cpp:16-57,
// tests/test-multiply.cpp:124-166: -m/-n/-k, --tile_{m,n,k}, --n_streams, -r/--n_rep, --ld_{a,b,c}, -t/--transpose, --alpha,
// --beta) plus --type {s,d,c,z} and --gpus N.  A flat table of named values; no third-party option parser.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

namespace cli {

struct Option {
    const char* short_name;  // "" when there is none
    const char* long_name;
    const char* fallback;
    const char* help;
    std::string text;
    bool given = false;
};

class Args {
public:
    explicit Args(std::vector<Option> table) : table_(std::move(table)) {
        for (Option& o : table_) o.text = o.fallback;
    }

    // returns false (after printing a message) when the command line is malformed; --help sets help_requested
    bool read(int argc, char** argv) {
        for (int i = 1; i < argc; ++i) {
            std::string tok = argv[i], inline_text;
            bool has_inline = false;
            if (tok == "-h" || tok == "--help") { help_requested = true; continue; }
            Option* hit = nullptr;
            if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
                std::string name = tok.substr(2);
                const size_t eq = name.find('=');
                if (eq != std::string::npos) { inline_text = name.substr(eq + 1); name.resize(eq); has_inline = true; }
                for (Option& o : table_) if (name == o.long_name) hit = &o;
            } else if (tok.size() >= 2 && tok[0] == '-') {
                for (Option& o : table_) if (o.short_name[0] && tok[1] == o.short_name[0]) hit = &o;
                if (tok.size() > 2) { inline_text = tok.substr(2); has_inline = true; }
            }
            if (!hit) { std::cerr << "[ERROR]: unknown option " << tok << " (see --help)" << std::endl; return false; }
            if (!has_inline) {
                if (i + 1 >= argc) { std::cerr << "[ERROR]: option " << tok << " needs a value" << std::endl; return false; }
                inline_text = argv[++i];
            }
            hit->text = inline_text;
            hit->given = true;
        }
        return true;
    }

    const std::string& text(const char* long_name) const {
        for (const Option& o : table_) if (std::string(long_name) == o.long_name) return o.text;
        std::cerr << "[ERROR]: internal: option " << long_name << " is not declared" << std::endl;
        std::exit(2);
    }
    long long integer(const char* long_name) const { return std::atoll(text(long_name).c_str()); }
    double real(const char* long_name) const { return std::atof(text(long_name).c_str()); }

    void usage(const char* program, const char* what) const {
        std::cout << what << "\nUsage:\n  " << program << " [OPTION...]\n\n";
        for (const Option& o : table_) {
            std::string left = std::string("  ") + (o.short_name[0] ? std::string("-") + o.short_name + ", " : std::string("    ")) + "--" + o.long_name + " arg";
            if (left.size() < 26) left.resize(26, ' ');
            std::cout << left << " " << o.help << " (default: " << o.fallback << ")\n";
        }
        std::cout << "  -h, --help               Print usage\n";
    }

    bool help_requested = false;

private:
    std::vector<Option> table_;
};

inline std::vector<Option> gemm_options(bool with_repetitions) {
    std::vector<Option> t = {
        {"m", "m_dim", "1000", "The number of rows of the resulting matrix C."},
        {"n", "n_dim", "1000", "The number of columns of the resulting matrix C."},
        {"k", "k_dim", "1000", "The size of the shared dimension between matrices A and B."},
        {"", "tile_m", "5000", "The tile size for dimension m (a staging hint: never changes results)."},
        {"", "tile_n", "5000", "The tile size for dimension n."},
        {"", "tile_k", "5000", "The tile size for dimension k."},
        {"", "n_streams", "2", "The number of GPU streams to use (hint)."},
        {"", "ld_a", "0", "The leading dimension of matrix A."},
        {"", "ld_b", "0", "The leading dimension of matrix B."},
        {"", "ld_c", "0", "The leading dimension of matrix C."},
        {"t", "transpose", "NN", "Two letters from {N,T,C}: op(A) and op(B); NT = A not transposed, B transposed."},
        {"", "alpha", "1.0", "The constant alpha in: C = beta*C + alpha*A*B."},
        {"", "beta", "0.0", "The constant beta in: C = beta*C + alpha*A*B."},
        {"", "type", "d", "Scalar type: s (float), d (double), c (complex<float>), z (complex<double>)."},
        {"", "gpus", "1", "Number of GPUs of this box to spread the C blocks over (copy-back runs only)."},
    };
    if (with_repetitions) t.insert(t.begin() + 7, {"r", "n_rep", "2", "The number of repetitions."});
    return t;
}

// "nt" -> {'N','T'}; false when not two letters of N/T/C
inline bool parse_transpose(std::string s, char* ta, char* tb) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char ch) { return (char)std::toupper(ch); });
    const std::string ok = "NTC";
    if (s.size() != 2 || ok.find(s[0]) == std::string::npos || ok.find(s[1]) == std::string::npos) return false;
    *ta = s[0]; *tb = s[1];
    return true;
}

struct Problem {
    long long m, n, k, tile_m, tile_n, tile_k, n_streams, ld_a, ld_b, ld_c, gpus;
    long long warm_size = 0;
    long long a_rows, a_cols, b_rows, b_cols;
    char trans_a, trans_b, type;
    double alpha, beta;
};

// false (message printed) when something is out of range
inline bool problem_from(const Args& a, Problem* p) {
    p->m = a.integer("m_dim"); p->n = a.integer("n_dim"); p->k = a.integer("k_dim");
    p->tile_m = a.integer("tile_m"); p->tile_n = a.integer("tile_n"); p->tile_k = a.integer("tile_k");
    p->n_streams = a.integer("n_streams"); p->gpus = a.integer("gpus");
    p->alpha = a.real("alpha"); p->beta = a.real("beta");
    if (!parse_transpose(a.text("transpose"), &p->trans_a, &p->trans_b)) {
        std::cout << "[ERROR]: --transpose option can only take two letters from N, T, C, e.g. NN, TN, NC, CT" << std::endl;
        return false;
    }
    const std::string& ty = a.text("type");
    p->type = ty.empty() ? 'd' : (char)std::tolower((unsigned char)ty[0]);
    if (std::string("sdcz").find(p->type) == std::string::npos) { std::cout << "[ERROR]: --type must be one of s, d, c, z" << std::endl; return false; }
    if (p->m < 1 || p->n < 1 || p->k < 1 || p->tile_m < 1 || p->tile_n < 1 || p->tile_k < 1 || p->n_streams < 1 || p->gpus < 1) {
        std::cout << "[ERROR]: dimensions, tile sizes, stream and GPU counts must be positive" << std::endl;
        return false;
    }
    p->a_rows = p->trans_a == 'N' ? p->m : p->k; p->a_cols = p->trans_a == 'N' ? p->k : p->m;
    p->b_rows = p->trans_b == 'N' ? p->k : p->n; p->b_cols = p->trans_b == 'N' ? p->n : p->k;
    p->ld_a = std::max(p->a_rows, a.integer("ld_a"));
    p->ld_b = std::max(p->b_rows, a.integer("ld_b"));
    p->ld_c = std::max(p->m, a.integer("ld_c"));
    return true;
}

// the banner both reference apps print (examples/multiply.cpp:119-153), same lines in the same order
inline void print_banner(const Problem& p, long long repetitions) {
    const char* bar = "=============================";
    std::cout << "==================================================\n"
              << "                Benchmarking Tiled-MM    \n"
              << "==================================================\n"
              << "         MATRIX SIZES \n" << bar << "\n"
              << " A = (" << p.a_rows << ", " << p.a_cols << ")\n"
              << " B = (" << p.b_rows << ", " << p.b_cols << ")\n"
              << " C = (" << p.m << ", " << p.n << ")\n" << bar << "\n"
              << "         LEADING DIMS \n" << bar << "\n"
              << " LD_A = " << p.ld_a << "\n LD_B = " << p.ld_b << "\n LD_C = " << p.ld_c << "\n" << bar << "\n"
              << "      SCALING CONSTANTS \n" << bar << "\n"
              << " alpha = " << p.alpha << "\n beta  = " << p.beta << "\n" << bar << "\n"
              << "      TRANSPOSE FLAGS \n" << bar << "\n"
              << " trans_a = " << p.trans_a << "\n trans_b = " << p.trans_b << "\n" << bar << "\n"
              << "         TILE SIZES \n" << bar << "\n"
              << " tile_m = " << p.tile_m << "\n tile_n = " << p.tile_n << "\n tile_k = " << p.tile_k << "\n" << bar << "\n"
              << "      ADDITIONAL OPTIONS \n" << bar << "\n"
              << " num. of gpu streams = " << p.n_streams << "\n num. of repetitions = " << repetitions << "\n"
              << " scalar type = " << p.type << "\n num. of gpus = " << p.gpus << "\n" << bar << std::endl;
}

}  // namespace cli
