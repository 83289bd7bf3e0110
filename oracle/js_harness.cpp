// js_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// End-to-end TurboPLONK join-split prover (BASELINE.json config #4) driven through the UNMODIFIED reference:
//   rollup::proofs::join_split::init_proving_key / init_verification_key   bb/rollup/proofs/join_split/join_split.cpp:15-41
//   rollup::proofs::join_split::noop_tx                                    bb/rollup/proofs/join_split/compute_circuit_data.cpp:28-59
//   rollup::proofs::join_split::new_join_split_prover + construct_proof    bb/rollup/proofs/join_split/join_split.cpp:50-63, prover.cpp:420-436
//   rollup::proofs::join_split::verify_proof                               bb/rollup/proofs/join_split/join_split.cpp:65-75
//
// oracle/Makefile links this file twice against the same reference objects:
//   oracle/_ref/js_prover_cpu   every symbol resolved inside the reference (the CPU prover)
//   oracle/_ref/js_prover_gpu   pippenger / pippenger_unsafe / Pippenger / fft / ifft / coset_fft ... weakened in the
//                               reference objects and resolved to aztec-2.0_b200/host/bbg_shim.cpp -> libbbg.so (CUDA)
// The only other change, identical in both binaries, makes the run reproducible: numeric::random::get_engine()
// (bb/numeric/random/engine.cpp:138-150, seeded from the OS) is weakened and defined here as a default-seeded
// engine, so the blinding scalars (prover.cpp:181-183, permutation_widget_impl.hpp:289-291) and noop_tx()'s
// keys are the same in both processes and the 1 952-byte proofs can be compared byte for byte.
//
// Output: one JSON object on stdout.
#include <chrono>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <numeric/random/engine.hpp>
#include <plonk/proof_system/commitment_scheme/kate_commitment_scheme.hpp>
#include <plonk/reference_string/file_reference_string.hpp>
#include <rollup/proofs/join_split/compute_circuit_data.hpp>
#include <rollup/proofs/join_split/join_split.hpp>

namespace numeric {
namespace random {
// deterministic replacement (see the header comment); Engine() is default-seeded std::mt19937_64 (engine.hpp)
Engine& get_engine()
{
    static Engine engine;
    engine.is_debug = true; // non-debug engines ignore their state and read std::random_device on every draw (engine.cpp:49-56)
    return engine;
}
} // namespace random
} // namespace numeric

namespace {
double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
std::string hex(const std::vector<uint8_t>& v)
{
    static const char* d = "0123456789abcdef";
    std::string s;
    s.reserve(v.size() * 2);
    for (uint8_t b : v) {
        s.push_back(d[b >> 4]);
        s.push_back(d[b & 15]);
    }
    return s;
}
} // namespace

int main(int argc, char** argv)
{
    using namespace rollup::proofs::join_split;
    const std::string srs = argc > 1 ? argv[1] : "../srs_db";
    const int reps = argc > 2 ? atoi(argv[2]) : 1;

    // present only in the binaries that link libbbg.so: create the CUDA context (driver initialisation + module load, a
    // one-off per process that has nothing to do with key generation) before the clock starts, and report it separately
    double t_cuda = 0.0;
    if (void* f = dlsym(RTLD_DEFAULT, "bbg_init")) {
        double tc = now();
        reinterpret_cast<int (*)(int)>(f)(-1);
        t_cuda = now() - tc;
    }

    // key generation exactly as the reference's own heavy test does it (join_split.test.cpp:44-50): proving key from the
    // circuit shape, verification key = 15 MSMs over the selector / permutation polynomials
    double t0 = now();
    init_proving_key(std::make_unique<waffle::FileReferenceStringFactory>(srs));
    init_verification_key(std::make_unique<waffle::FileReferenceStringFactory>(srs));
    double t_keys = now() - t0;
    // once more with everything warm (CUDA kernels loaded, library workspaces and tables allocated): what key generation
    // costs a process that has generated a key before
    t0 = now();
    init_proving_key(std::make_unique<waffle::FileReferenceStringFactory>(srs));
    init_verification_key(std::make_unique<waffle::FileReferenceStringFactory>(srs));
    double t_keys_warm = now() - t0;

    // PCIe bytes per proof, when libbbg is underneath (bbg_stats_totals: calls / H2D / D2H of msm, ntt, srs, poly)
    typedef int (*totals_fn)(uint64_t*);
    totals_fn totals = reinterpret_cast<totals_fn>(dlsym(RTLD_DEFAULT, "bbg_stats_totals"));
    auto pcie_now = [&](uint64_t& h2d, uint64_t& d2h) {
        h2d = d2h = 0;
        if (!totals) return;
        uint64_t t[12];
        totals(t);
        for (int i = 0; i < 4; ++i) {
            h2d += t[3 * i + 1];
            d2h += t[3 * i + 2];
        }
    };
    std::vector<uint8_t> proof, first_proof;
    std::string times;
    bool ok = true;
    size_t gates = 0;
    for (int r = 0; r < reps; ++r) {
        join_split_tx tx = noop_tx(); // deterministic: get_engine() above
        double t1 = now();
        auto prover = new_join_split_prover(tx); // circuit + witness (CPU in both binaries)
        double t2 = now();
        uint64_t h0, d0, h1, d1;
        pcie_now(h0, d0);
        proof = prover.construct_proof().proof_data; // the part the hot path accelerates
        double t3 = now();
        pcie_now(h1, d1);
        gates = prover.get_circuit_size();
        times += (r ? ", " : "") + std::string("{\"witness_s\": ") + std::to_string(t2 - t1) + ", \"construct_proof_s\": " +
                 std::to_string(t3 - t2) + ", \"h2d_bytes\": " + std::to_string(h1 - h0) + ", \"d2h_bytes\": " + std::to_string(d1 - d0) + "}";
        ok = ok && verify_proof(waffle::plonk_proof{ proof });
        if (r == 0) first_proof = proof;
    }

    // present only in the binary that links libbbg.so: how many CUDA kernels the prover's hot path launched
    unsigned long long launches = 0;
    if (void* f = dlsym(RTLD_DEFAULT, "bbg_kernel_launches")) {
        launches = reinterpret_cast<unsigned long long (*)()>(f)();
    }
    printf("{\"gpu_kernel_launches\": %llu, \"cuda_init_s\": %.6f, ", launches, t_cuda);
    printf("\"keygen_warm_s\": %.6f, ", t_keys_warm);
    printf("\"n\": %zu, \"proof_bytes\": %zu, \"keygen_s\": %.6f, \"proofs\": [%s], \"verified\": %s, "
           "\"first_proof\": \"%s\", \"last_proof\": \"%s\"}\n",
           gates, proof.size(), t_keys, times.c_str(), ok ? "true" : "false", hex(first_proof).c_str(), hex(proof).c_str());
    return ok ? 0 : 1;
}
