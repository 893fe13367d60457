// Stand-in for alexstraub1990/simultaneous-sort (vcpkg port, absent here). OUR code, not reference code.
// Only ParticlesToDensity's vector mode (aggregator 2) calls sort_with; it is never on the parity path.
// Semantics: sort the first container with `comp` and apply the same permutation to all the others.
#pragma once
#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>
template<class Comp, class V0, class... Vs>
void sort_with(Comp comp, V0& keys, Vs&... others) {
    std::vector<std::size_t> perm(keys.size());
    std::iota(perm.begin(), perm.end(), std::size_t{0});
    std::stable_sort(perm.begin(), perm.end(), [&](std::size_t a, std::size_t b) { return comp(keys[a], keys[b]); });
    auto apply = [&](auto& v) {
        auto copy = v;
        for (std::size_t i = 0; i < perm.size(); ++i) v[i] = copy[perm[i]];
    };
    apply(keys);
    (apply(others), ...);
}
