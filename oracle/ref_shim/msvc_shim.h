// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
// Force-included (-include) in front of the UNMODIFIED reference
// SeSchwarzPreconditioner.cpp so that its MSVC-only `#ifdef WIN32` AVX2 bodies
// (the real LDLtInverse512 / SchwarzLocalXSym, cpp:1394-1469, 1622-1655)
// compile with g++.  Nothing here restates reference logic; it only maps
// MSVC spellings onto GCC builtins.
#pragma once
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <memory>
#include <vector>
#include <algorithm>
#include <immintrin.h>

using std::isnan;
using std::isinf;

#define __int64 long long
#define __pragma(x) _Pragma(#x)

static inline unsigned int __lzcnt(unsigned int v) { return v ? (unsigned)__builtin_clz(v) : 32u; }
static inline unsigned int __popcnt(unsigned int v) { return (unsigned)__builtin_popcount(v); }
static inline unsigned long long __popcnt64(unsigned long long v) { return (unsigned long long)__builtin_popcountll(v); }
static inline unsigned char _BitScanForward(unsigned long* index, unsigned long mask)
{
	// SeIntrinsic.h:54-60 passes an int; only the low 32 bits are meaningful.
	unsigned int m = (unsigned int)mask;
	if (!m) return 0;
	*index = (unsigned long)__builtin_ctz(m);
	return 1;
}
