// Host build of aztec-2.0_b200/csrc/inv.cuh (the same source the CUDA kernels compile) for tests/test_host_inverse.py.
#include "../../aztec-2.0_b200/csrc/inv.cuh"

extern "C" {
// out = a^-1 2^k mod p, returns k
unsigned inv_almost(const uint32_t* a, const uint32_t* p, uint32_t* out)
{
    uint32_t aa[8], pp[8], oo[8];
    for (int i = 0; i < 8; ++i) { aa[i] = a[i]; pp[i] = p[i]; }
    bbg::inv::Kaliski st;
    st.init(aa, pp);
    while (st.alive()) st.step();
    st.step(); // a finished lane must be left untouched by further steps (other lanes of its warp keep going)
    st.step();
    st.finish(oo, pp);
    for (int i = 0; i < 8; ++i) out[i] = oo[i];
    return st.k;
}
void inv_fix_table(const uint32_t* p, const uint32_t* r2, uint32_t* table)
{
    uint32_t pp[8], rr[8];
    for (int i = 0; i < 8; ++i) { pp[i] = p[i]; rr[i] = r2[i]; }
    bbg::inv::inv_fix_table(pp, rr, table);
}
unsigned inv_fix_entries() { return bbg::inv::FIX_ENTRIES; }
unsigned inv_k_min() { return bbg::inv::K_MIN; }
}
