// Test helper: the product's restated libstdc++ nth_element (eventcalib_b200/csrc/ecb_nth_element.h) next to the real
// std::nth_element on the same input; both permute a copy, the caller compares the whole arrays.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../eventcalib_b200/csrc/ecb_nth_element.h"

extern "C" {
// v: values to permute (ids), key[id]: sort key.  out_std / out_ecb: permuted copies.
void nth_both(const uint32_t *v, int n, const uint32_t *key, int nth, uint32_t *out_std, uint32_t *out_ecb) {
    std::vector<uint32_t> a(v, v + n), b(v, v + n);
    auto less = [&](uint32_t l, uint32_t r) { return key[l] < key[r]; };
    std::nth_element(a.begin(), a.begin() + nth, a.end(), less);
    ecb_nth::nth_element(b.data(), (long) n, (long) nth, less);
    std::copy(a.begin(), a.end(), out_std);
    std::copy(b.begin(), b.end(), out_ecb);
}
}
