// Host-side unit test shim for csrc/staging.hpp's CopyPool (the thread pool behind the pageable-memory staging path),
// compiled with g++ from the very header the library includes.  TEST INFRASTRUCTURE ONLY.
#include "../../aztec-2.0_b200/csrc/staging.hpp"

namespace bbg {
void set_last_error(const std::string&) {}
} // namespace bbg

extern "C" {
void* pool_new(unsigned threads) { return new bbg::CopyPool(threads); }
void pool_delete(void* p) { delete reinterpret_cast<bbg::CopyPool*>(p); }
void pool_copy(void* p, void* dst, const void* src, size_t bytes) { reinterpret_cast<bbg::CopyPool*>(p)->copy(dst, src, bytes); }
}
