// TEST INFRASTRUCTURE ONLY.  One CUDA thread block played by OS threads, for running device code of csrc/*.cuh on the CPU:
// every CUDA thread is a std::thread, __syncthreads / __syncwarp / named warp barriers are std::barrier objects, warp shuffles
// and the m16n8k8 TF32 MMA exchange their operands through per-warp scratch (fragment layouts of the PTX ISA), TF32
// conversion rounds to 10 mantissa bits (ties away from zero, like cvt.rna).  Arithmetic intrinsics map to fmaf & co., so
// FP32 results equal the GPU's wherever the GPU code uses the IEEE intrinsics.  Data races are NOT detected; this checks
// logic and indexing, not synchronisation.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>   // float4, make_float4, the __device__ / __forceinline__ decorations as no-ops for g++

namespace emu {

struct Dim { unsigned x = 0, y = 0, z = 0; };
inline thread_local Dim tid;
inline Dim bdim, bid, gdim;

struct Block
{
	int nThreads;
	std::barrier<> all;
	std::vector<std::unique_ptr<std::barrier<>>> warp;
	std::vector<float> shfl;            // [warp][lane]
	std::vector<unsigned> shflu;        // [warp][lane]
	std::vector<unsigned> fragA, fragB; // [warp][lane][4], [warp][lane][2]
	explicit Block(int n) : nThreads(n), all(n), shfl((size_t)n), shflu((size_t)n), fragA((size_t)n * 4), fragB((size_t)n * 2)
	{
		for (int w = 0; w < n / 32; ++w) warp.emplace_back(new std::barrier<>(32));
	}
};
inline Block* block = nullptr;

// A whole grid, one block after the other (kernels without inter-block synchronisation).  The OS threads are created once
// per launch and walk through the blocks together: `edge` separates consecutive blocks (a thread that returns early from the
// kernel waits there), `all` is the kernel's own __syncthreads.
inline void launch(int grid, int nThreads, const std::function<void()>& kernel)
{
	Block b(nThreads);
	std::barrier<> edge(nThreads);
	block = &b;
	bdim.x = (unsigned)nThreads; bdim.y = bdim.z = 1;
	gdim.x = (unsigned)grid; gdim.y = gdim.z = 1;
	std::vector<std::thread> th;
	for (int t = 0; t < nThreads; ++t)
		th.emplace_back([t, grid, &kernel, &edge] {
			tid.x = (unsigned)t;
			for (int blk = 0; blk < grid; ++blk)
			{
				if (t == 0) bid.x = (unsigned)blk;
				edge.arrive_and_wait();
				kernel();
				edge.arrive_and_wait();
			}
		});
	for (auto& x : th) x.join();
	block = nullptr;
}

inline void run(int nThreads, const std::function<void()>& kernel) { launch(1, nThreads, kernel); }

}  // namespace emu

#define blockIdx (::emu::bid)
#define gridDim (::emu::gdim)
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __global__
#define __global__
#undef __shared__
#define __shared__ static      // one copy for all threads of the (single, sequentially emulated) block
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#define threadIdx (::emu::tid)
#define blockDim (::emu::bdim)

inline void __syncthreads() { emu::block->all.arrive_and_wait(); }
inline void emu_warp_barrier() { emu::block->warp[emu::tid.x >> 5]->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier(); }

inline float __shfl_sync(unsigned, float v, int src)
{
	const unsigned w = emu::tid.x >> 5;
	emu::block->shfl[emu::tid.x] = v;
	emu_warp_barrier();
	const float r = emu::block->shfl[32 * w + (src & 31)];
	emu_warp_barrier();
	return r;
}

inline unsigned emu_exchange(unsigned v, int src)
{
	const unsigned w = emu::tid.x >> 5;
	emu::block->shflu[emu::tid.x] = v;
	emu_warp_barrier();
	const unsigned r = emu::block->shflu[32 * w + (src & 31)];
	emu_warp_barrier();
	return r;
}
inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return emu_exchange(v, src); }
inline int __shfl_sync(unsigned, int v, int src) { return (int)emu_exchange((unsigned)v, src); }
inline float __shfl_down_sync(unsigned m, float v, int delta)
{
	const int lane = (int)(emu::tid.x & 31);
	return __shfl_sync(m, v, lane + delta < 32 ? lane + delta : lane);
}
inline float __shfl_xor_sync(unsigned m, float v, int mask) { return __shfl_sync(m, v, (int)((emu::tid.x & 31) ^ (unsigned)mask)); }
inline unsigned __ballot_sync(unsigned, int pred)
{
	const unsigned w = emu::tid.x >> 5;
	emu::block->shflu[emu::tid.x] = pred ? 1u : 0u;
	emu_warp_barrier();
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) r |= emu::block->shflu[32 * w + l] << l;
	emu_warp_barrier();
	return r;
}
inline unsigned __match_any_sync(unsigned, int key)
{
	const unsigned w = emu::tid.x >> 5;
	emu::block->shflu[emu::tid.x] = (unsigned)key;
	emu_warp_barrier();
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) r |= (emu::block->shflu[32 * w + l] == (unsigned)key ? 1u : 0u) << l;
	emu_warp_barrier();
	return r;
}
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __shfl_up_sync(unsigned, int v, int delta)
{
	const int lane = (int)(emu::tid.x & 31);
	const int r = (int)emu_exchange((unsigned)v, lane >= delta ? lane - delta : lane);
	return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)      // every lane passes the mask of its own group
{
	const unsigned w = emu::tid.x >> 5;
	emu::block->shflu[emu::tid.x] = v;
	emu_warp_barrier();
	unsigned r = 0;
	for (int l = 0; l < 32; ++l)
		if ((mask >> l) & 1u) r |= emu::block->shflu[32 * w + l];
	emu_warp_barrier();
	return r;
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v)
{
	unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
	return old;
}
inline int __shfl_sync(unsigned m, unsigned long long, int) = delete;
inline double __shfl_xor_sync(unsigned m, double v, int mask)
{
	unsigned long long u;
	std::memcpy(&u, &v, 8);
	const int src = (int)((emu::tid.x & 31) ^ (unsigned)mask);
	const unsigned lo = emu_exchange((unsigned)u, src), hi = emu_exchange((unsigned)(u >> 32), src);
	u = ((unsigned long long)hi << 32) | lo;
	std::memcpy(&v, &u, 8);
	return v;
}
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }

// atomics: CUDA threads of one block run concurrently here, so these are real atomics (compare-exchange loops)
template <typename T>
inline T emu_atomic_add(T* p, T v)
{
	T old;
	__atomic_load(p, &old, __ATOMIC_RELAXED);
	for (;;)
	{
		T want = old + v;
		if (__atomic_compare_exchange(p, &old, &want, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return old;
	}
}
inline double atomicAdd(double* p, double v) { return emu_atomic_add(p, v); }
inline float atomicAdd(float* p, float v) { return emu_atomic_add(p, v); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// dynamic shared memory of the block being emulated (blocks run one after the other)
inline unsigned char* emu_dynamic_smem()
{
	alignas(128) static unsigned char buf[128 * 1024];
	return buf;
}
inline void __threadfence_block() {}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline int __shfl_xor_sync(unsigned m, int v, int mask) { return __shfl_sync(m, v, (int)((emu::tid.x & 31) ^ (unsigned)mask)); }

inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

// cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits; low 13 bits cleared
inline unsigned emu_cvt_rna_tf32(float x)
{
	unsigned u = __float_as_uint(x);
	if ((u & 0x7f800000u) == 0x7f800000u) return u;
	u += 0x1000u;
	return u & ~0x1fffu;
}

// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32: D = A B + D over one warp.
//   A (16x8, row): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)          g = lane / 4, t = lane % 4
//   B (8x8, col) : b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   C/D (16x8)   : c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
inline void emu_mma_m16n8k8_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
	const unsigned w = emu::tid.x >> 5, lane = emu::tid.x & 31, g = lane >> 2, t = lane & 3;
	unsigned* A = emu::block->fragA.data() + (size_t)w * 32 * 4;
	unsigned* B = emu::block->fragB.data() + (size_t)w * 32 * 2;
	for (int u = 0; u < 4; ++u) A[4 * lane + u] = a[u];
	for (int u = 0; u < 2; ++u) B[2 * lane + u] = b[u];
	emu_warp_barrier();
	auto Ael = [&](unsigned row, unsigned k) { return __uint_as_float(A[4 * ((row & 7) * 4 + (k & 3)) + (row >> 3) + 2 * (k >> 2)]); };
	auto Bel = [&](unsigned k, unsigned n) { return __uint_as_float(B[2 * (n * 4 + (k & 3)) + (k >> 2)]); };
	for (int u = 0; u < 4; ++u)
	{
		const unsigned row = g + 8 * (u >> 1), col = 2 * t + (u & 1);
		float acc = d[u];
		for (unsigned k = 0; k < 8; ++k) acc = std::fmaf(Ael(row, k), Bel(k, col), acc);
		d[u] = acc;
	}
	emu_warp_barrier();
}
