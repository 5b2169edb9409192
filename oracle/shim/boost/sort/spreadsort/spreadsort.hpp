// Test-infrastructure shim (NOT product code): the reference's docidupdates.cpp includes
// boost's spreadsort, which is absent in this image. Only pack_updates() uses it (off the
// hot path); std::sort is semantically identical for that use.
#pragma once
#include <algorithm>
namespace boost { namespace sort { namespace spreadsort {
template <class It> inline void spreadsort(It a, It b) { std::sort(a, b); }
}}}
