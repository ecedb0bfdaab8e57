// Per-PCG-iteration apply: z = sum_levels P_l * blockdiag(A_l)^-1 * P_l^T r.
// Replaces Preconditioning (SeSchwarzPreconditioner.cpp:100-110): the two memsets (102-103),
// BuildResidualHierarchy (1548-1598), SchwarzLocalXSym (1600-1696) and CollectFinalZ (1698-1719).
//
// Kernels:
//   restrict_fine   r (original order) -> level-1 residuals; level 0 is never materialised
//   restrict_l1     level 1 -> 2 (and 2 -> 3 on large meshes);  restrict_top: one CTA walks the top levels
//   gather_peers    (sharded contexts) publish / wait / pull of the other ranks' coarse residuals over peer memory
//   solve_coarse    Z_l = inv_l * R_l for every block of levels >= 1, four warps per block (latency-bound)
//   prolong_sum     per level-1 node: Z_1 + Z_2[parent] + ... (what CollectFinalZ adds to each of its vertices)
//   solve_fine      gathers r again (L2 resident), multiplies by the packed level-0 inverse, adds the level-1 sum
//                   and scatters z straight to original order
//   add_coarse      z += level-1 sum, for the fine banks that were solved before the coarse levels finished
// Captured once as a CUDA graph with two branches: the coarse chain (restrict_fine -> restrict_l1 -> restrict_top -> solve_coarse -> prolong_sum, ~3 % of
// the bytes but latency-bound) runs CONCURRENTLY with solve_fine over the first part of the fine banks (which cannot add
// the coarse part yet; add_coarse does that afterwards), then solve_fine with the fused addition runs over the rest.
// solve_fine moves 97 % of the bytes: 18,624 B of packed inverse per 32 vertices against 32 x (16+16) B of r/z,
// i.e. ~1 FLOP per byte, HBM-bound.  One warp owns one domain; lane i owns node i (3 rows).  The packed layout
// (mas_internal.h) makes every load a fully coalesced 512-byte LDG.128 per warp; the symmetric half of each
// 3x3 block is applied through two warp shuffles (x of the column node in, B^T x back out), so each matrix
// element is read exactly once from HBM and never staged.
//
// Every sum is evaluated in a fixed order (no float atomics): results are run-to-run deterministic.
#include "mas_internal.h"

namespace mas {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kApplyThreads = 256;
constexpr int kWarpsPerCta = kApplyThreads / 32;

// streaming 128-bit load: read-only path, no L1 allocation (each byte is used once)
#ifdef MAS_CPU_EMULATION   // tests/emu/apply_emu.cpp compiles the kernels of this file for the host (test infrastructure)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) { return *p; }
__device__ __forceinline__ float ldg_stream1(const float* p) { return *p; }
#else
__device__ __forceinline__ float4 ldg_stream4(const float4* p)
{
	float4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p)
{
	float v;
	asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
	return v;
}
#endif

struct Vec3
{
	float x, y, z;
};

// one full cyclic block diagonal d: lane i holds B = A(i, (i+d)&31) row-major in b[0..8]
__device__ __forceinline__ void apply_block(const float* b, int d, int lane, const Vec3& x, Vec3& y)
{
	const int src = (lane + d) & 31, dst = (lane - d) & 31;
	const float xjx = __shfl_sync(kFull, x.x, src), xjy = __shfl_sync(kFull, x.y, src), xjz = __shfl_sync(kFull, x.z, src);
	y.x = fmaf(b[0], xjx, fmaf(b[1], xjy, fmaf(b[2], xjz, y.x)));
	y.y = fmaf(b[3], xjx, fmaf(b[4], xjy, fmaf(b[5], xjz, y.y)));
	y.z = fmaf(b[6], xjx, fmaf(b[7], xjy, fmaf(b[8], xjz, y.z)));
	const float tx = fmaf(b[0], x.x, fmaf(b[3], x.y, b[6] * x.z));
	const float ty = fmaf(b[1], x.x, fmaf(b[4], x.y, b[7] * x.z));
	const float tz = fmaf(b[2], x.x, fmaf(b[5], x.y, b[8] * x.z));
	y.x += __shfl_sync(kFull, tx, dst);
	y.y += __shfl_sync(kFull, ty, dst);
	y.z += __shfl_sync(kFull, tz, dst);
}

// y = (packed symmetric 96x96 block) * x for one domain, one warp, direct coalesced LDG.128 stream
__device__ __forceinline__ Vec3 solve_domain_ldg(const float* __restrict__ blk, int lane, const Vec3 x)
{
	const float4* p4 = reinterpret_cast<const float4*>(blk) + lane;
	Vec3 y = { 0.f, 0.f, 0.f };
	float4 bufA[9], bufB[9];
#pragma unroll
	for (int q = 0; q < 9; ++q) bufA[q] = ldg_stream4(p4 + 32 * q);
#pragma unroll
	for (int q = 0; q < 9; ++q) bufB[q] = ldg_stream4(p4 + 32 * (9 + q));
	{
		const float* m = reinterpret_cast<const float*>(bufA);
#pragma unroll
		for (int dd = 0; dd < 4; ++dd) apply_block(m + 9 * dd, 1 + dd, lane, x, y);
	}
#pragma unroll
	for (int q = 0; q < 9; ++q) bufA[q] = ldg_stream4(p4 + 32 * (18 + q));
	{
		const float* m = reinterpret_cast<const float*>(bufB);
#pragma unroll
		for (int dd = 0; dd < 4; ++dd) apply_block(m + 9 * dd, 5 + dd, lane, x, y);
	}
	// remainder: slots 27..34 (d = 13,14,15 and 5 of the 6 diagonal floats), the tail float, the half diagonal
#pragma unroll
	for (int q = 0; q < 8; ++q) bufB[q] = ldg_stream4(p4 + 32 * (27 + q));
	const float dTail = ldg_stream1(blk + kTailBase + lane);
	float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
	float h8 = 0.f;
	if (lane < 16)
	{
		const float4* ph = reinterpret_cast<const float4*>(blk + kHalfBase) + lane;
		h0 = ldg_stream4(ph);
		h1 = ldg_stream4(ph + 16);
		h8 = ldg_stream1(blk + kHalfTail + lane);
	}
	{
		const float* m = reinterpret_cast<const float*>(bufA);
#pragma unroll
		for (int dd = 0; dd < 4; ++dd) apply_block(m + 9 * dd, 9 + dd, lane, x, y);
	}
	{
		const float* m = reinterpret_cast<const float*>(bufB);
#pragma unroll
		for (int dd = 0; dd < 3; ++dd) apply_block(m + 9 * dd, 13 + dd, lane, x, y);
		// diagonal block, 6 unique floats: (0,0) (1,0) (1,1) (2,0) (2,1) | (2,2) is the tail float
		const float d0 = m[27], d1 = m[28], d2 = m[29], d3 = m[30], d4 = m[31], d5 = dTail;
		y.x = fmaf(d0, x.x, fmaf(d1, x.y, fmaf(d3, x.z, y.x)));
		y.y = fmaf(d1, x.x, fmaf(d2, x.y, fmaf(d4, x.z, y.y)));
		y.z = fmaf(d3, x.x, fmaf(d4, x.y, fmaf(d5, x.z, y.z)));
	}
	{
		// half diagonal: lanes i < 16 hold B = A(i, i+16)
		const float b[9] = { h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h8 };
		const int peer = lane ^ 16;
		const float xjx = __shfl_sync(kFull, x.x, peer), xjy = __shfl_sync(kFull, x.y, peer), xjz = __shfl_sync(kFull, x.z, peer);
		float tx = 0.f, ty = 0.f, tz = 0.f;
		if (lane < 16)
		{
			y.x = fmaf(b[0], xjx, fmaf(b[1], xjy, fmaf(b[2], xjz, y.x)));
			y.y = fmaf(b[3], xjx, fmaf(b[4], xjy, fmaf(b[5], xjz, y.y)));
			y.z = fmaf(b[6], xjx, fmaf(b[7], xjy, fmaf(b[8], xjz, y.z)));
			tx = fmaf(b[0], x.x, fmaf(b[3], x.y, b[6] * x.z));
			ty = fmaf(b[1], x.x, fmaf(b[4], x.y, b[7] * x.z));
			tz = fmaf(b[2], x.x, fmaf(b[5], x.y, b[8] * x.z));
		}
		const float rx = __shfl_sync(kFull, tx, peer), ry = __shfl_sync(kFull, ty, peer), rz = __shfl_sync(kFull, tz, peer);
		if (lane >= 16) { y.x += rx; y.y += ry; y.z += rz; }
	}
	return y;
}

// Sum `val` over the lanes that share `key` (key < 0: lane takes no part) in a fixed butterfly order and let the
// lowest lane of each group store it.  All children of a coarse node sit in one bank (clustering is per bank),
// so a plain store is enough: no atomics, deterministic.
__device__ __forceinline__ void group_sum_store(int key, Vec3 val, int lane, float4* __restrict__ out, int outBase,
	float4* __restrict__ out2 = nullptr)
{
	unsigned peers = __match_any_sync(kFull, key);
	unsigned todo = __ballot_sync(kFull, key >= 0 && lane == __ffs(peers) - 1);  // one bit per group (its leader)
	while (todo)
	{
		const int leader = __ffs(todo) - 1;
		todo &= todo - 1;
		const unsigned grp = __shfl_sync(kFull, peers, leader);
		const bool in = (grp >> lane) & 1u;
		float sx = in ? val.x : 0.f, sy = in ? val.y : 0.f, sz = in ? val.z : 0.f;
#pragma unroll
		for (int off = 16; off > 0; off >>= 1)
		{
			sx += __shfl_xor_sync(kFull, sx, off);
			sy += __shfl_xor_sync(kFull, sy, off);
			sz += __shfl_xor_sync(kFull, sz, off);
		}
		if (lane == leader)
		{
			out[key - outBase] = make_float4(sx, sy, sz, 0.f);
			if (out2) out2[key - outBase] = make_float4(sx, sy, sz, 0.f);
		}
	}
}

// BuildResidualHierarchy, level 0 -> 1 (cpp:1558-1574).  `send` (sharded contexts with attached peers): second copy of the
// results in this rank's peer-readable send buffer, selected by the parity of the next apply counter.
// The kernel is a chain of dependent loads (s2o -> r) per bank, so every warp carries kRestrictBanks banks with all their
// loads in flight together: it heads the latency-bound coarse chain and shares the SMs with the streaming fine solve.
constexpr int kRestrictBanks = 4;
__device__ __forceinline__ void restrict_fine_warp(const int bank0, const int lane, const float4* __restrict__ r, const int* __restrict__ s2o,
	const int* __restrict__ goingNext, int nv, int nVC, int bankEnd, float4* __restrict__ coarseR, float4* __restrict__ second)
{
	if (bank0 >= bankEnd) return;
	int ov[kRestrictBanks], key[kRestrictBanks];
	float4 rv[kRestrictBanks];
#pragma unroll
	for (int k = 0; k < kRestrictBanks; ++k)
	{
		const int v = (bank0 + k) * 32 + lane;
		const bool live = bank0 + k < bankEnd && v < nv;
		ov[k] = live ? s2o[v] : -1;
		key[k] = live ? goingNext[v] : -1;
	}
#pragma unroll
	for (int k = 0; k < kRestrictBanks; ++k) rv[k] = ov[k] >= 0 ? r[ov[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
	for (int k = 0; k < kRestrictBanks; ++k)
	{
		if (bank0 + k >= bankEnd) break;
		const Vec3 val = { rv[k].x, rv[k].y, rv[k].z };
		group_sum_store(key[k], val, lane, coarseR, nVC, second);
	}
}

__global__ void __launch_bounds__(kApplyThreads) restrict_fine_kernel(const float4* __restrict__ r, const int* __restrict__ s2o,
	const int* __restrict__ goingNext, int nv, int nVC, int bankBegin, int bankEnd, float4* __restrict__ coarseR,
	float4* __restrict__ send, unsigned long long sendCap, const unsigned* epoch)
{
	const int lane = threadIdx.x & 31;
	const int bank0 = bankBegin + (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * kRestrictBanks;
	if (bank0 >= bankEnd) return;
	float4* second = nullptr;
	if (send) second = send + (unsigned long long)((*reinterpret_cast<const volatile unsigned*>(epoch) + 1u) & 1u) * sendCap;
	restrict_fine_warp(bank0, lane, r, s2o, goingNext, nv, nVC, bankEnd, coarseR, second);
}

// ---- multi-GPU exchange over peer memory (NVLink) ---------------------------------------------------------------
// Every rank owns an "arena" that its peers have mapped: two SEND buffers for its level-1 residuals (double-buffered by
// the parity of the apply counter) and one arrival flag per rank.  The level-1 nodes a rank produces form one contiguous
// range (ids follow the Morton order of the fine banks) and have no other producer, so the exchange is an all-gather
// without a reduction, done as a PULL:
//   restrict_fine    writes its residuals to the local coarseR and to the local send buffer (plain local stores);
//   gather_peers     stores the apply counter into every rank's flag slot (the payload was completed by the previous
//                    kernel, so the flag store itself carries no fence), waits (load-acquire) until all ranks have
//                    published this apply, then reads the foreign slices straight out of the peers' send buffers over
//                    NVLink into the local coarseR.
// No NCCL call, no host involvement, no remote stores of payload (a push variant needed a system-scope fence in every
// storing CTA and cost 15-29 us per apply): the whole sharded apply stays one CUDA graph per rank.
// Why two buffers are enough: a peer can only be one apply ahead of a rank it still has to hear from.
constexpr int kMaxWorld = 16;
struct PeerArgs
{
	float4* send[kMaxWorld];      // arena base of every rank (send buffer b at send[q] + b * cap)
	unsigned* flags[kMaxWorld];   // flags array of every rank; slot [rank] is written by `rank`
	int sliceBegin[kMaxWorld + 1];// level-1 slice (coarse index) produced by rank q: [sliceBegin[q], sliceBegin[q+1])
	unsigned* epoch;              // local apply counter (number of completed exchanges)
	unsigned* ticket;             // local: blocks of gather_peers that have finished
	unsigned* error;              // local sticky error flag (peer wait timed out)
	unsigned long long cap;       // float4 elements per send buffer
	int world, rank;
	int strictPublish;            // MAS_OPT_STRICT_PUBLISH: system-scope fence before the flag stores
};

#ifndef MAS_CPU_EMULATION   // the peer exchange needs several GPUs: not part of the host emulation
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// Publish, wait, pull.  The residuals of this apply are complete in the local send buffer: they were written by the
// PREVIOUS kernel on this stream, so they already sit in this GPU's L2 (the point of coherence for reads arriving over
// NVLink).  The flag store therefore needs no fence of its own (a release store here costs a MEMBAR.SYS, ~5 us while the
// fine solve is streaming).  Block 0 stores this apply's number into every rank's flag slot; every block then waits until
// all ranks have published it and pulls its part of the foreign level-1 slices into coarseR.  The last block to finish
// advances the local apply counter (all blocks have read it by then).
constexpr int kGatherPerThread = 4;
__global__ void __launch_bounds__(256) gather_peers_kernel(PeerArgs pa, int first, int count, float4* __restrict__ coarseR)
{
	__shared__ int ok;
	const unsigned want = *reinterpret_cast<volatile unsigned*>(pa.epoch) + 1u;
	if (blockIdx.x == 0 && (int)threadIdx.x < pa.world)
	{
		// strict mode: order the payload (written by the previous kernel) before the flag at SYSTEM scope, as the PTX memory
		// model asks for between GPUs; the default relies on the kernel boundary having put the payload into this GPU's L2
		if (pa.strictPublish) __threadfence_system();
		asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(pa.flags[threadIdx.x] + pa.rank), "r"(want) : "memory");
	}
	if (threadIdx.x == 0) ok = 1;
	__syncthreads();
	if ((int)threadIdx.x < pa.world)
	{
		const unsigned* f = pa.flags[pa.rank] + threadIdx.x;
		const long long t0 = clock64();
		// counters only grow; a peer that is already one apply ahead satisfies this too (its data sits in the other buffer)
		while ((int)(ld_acquire_sys(f) - want) < 0)
		{
			if (clock64() - t0 > 4000000000ll)   // ~2 s: a rank is gone.  Sticky flag in host memory: every later call fails.
			{
				ok = 0;
				asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(pa.error), "r"(1u) : "memory");
				break;
			}
			__nanosleep(32);
		}
	}
	__syncthreads();
	if (ok)
	{
		const unsigned long long off = (unsigned long long)(want & 1u) * pa.cap;
		const int stride = gridDim.x * blockDim.x;
		for (int base = blockIdx.x * blockDim.x + threadIdx.x; base < count; base += kGatherPerThread * stride)
		{
			// all remote loads of a thread are issued before the first one is consumed: one NVLink round trip, not one per element
			const float4* src[kGatherPerThread];
			float4 v[kGatherPerThread];
#pragma unroll
			for (int k = 0; k < kGatherPerThread; ++k)
			{
				const int i = first + base + k * stride;
				src[k] = nullptr;
				if (base + k * stride < count)
				{
					int q = 0;
					while (q + 1 < pa.world && i >= pa.sliceBegin[q + 1]) ++q;
					if (q != pa.rank) src[k] = pa.send[q] + off + i;
				}
			}
#pragma unroll
			for (int k = 0; k < kGatherPerThread; ++k)
				if (src[k]) v[k] = __ldcg(src[k]);   // written by another GPU: read at the owner's L2, never from this SM's L1
#pragma unroll
			for (int k = 0; k < kGatherPerThread; ++k)
				if (src[k]) coarseR[first + base + k * stride] = v[k];
		}
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		__threadfence();
		if (atomicAdd(pa.ticket, 1u) == gridDim.x - 1)
		{
			*reinterpret_cast<volatile unsigned*>(pa.ticket) = 0u;
			*reinterpret_cast<volatile unsigned*>(pa.epoch) = want;
		}
	}
}

#endif  // MAS_CPU_EMULATION

// One 32-node group of coarse nodes [begin + 32*bank, ...): sum R over the nodes that share a parent and store it
// (BuildResidualHierarchy level l -> l+1, cpp:1577-1591)
__device__ __forceinline__ void restrict_bank(const int* __restrict__ goingNext, int begin, int count, int bank, int nVC,
	float4* __restrict__ coarseR, int lane, const float4 rv, float4* __restrict__ second = nullptr)
{
	const int local = bank * 32 + lane;
	int key = -1;
	Vec3 val = { 0.f, 0.f, 0.f };
	if (local < count)
	{
		val.x = rv.x; val.y = rv.y; val.z = rv.z;
		key = goingNext[begin + local];
	}
	group_sum_store(key, val, lane, coarseR, nVC, second);
}

// BuildResidualHierarchy, level 1 -> 2 (cpp:1577-1591): one warp per 32 level-1 nodes, banks [bankBegin, bankEnd).
// `send`: see restrict_fine_kernel (sharded contexts with aligned cuts publish their level-2 residuals instead).
__global__ void __launch_bounds__(kApplyThreads) restrict_l1_kernel(const int* __restrict__ goingNext, int begin, int count,
	int nVC, int bankBegin, int bankEnd, float4* __restrict__ coarseR, float4* __restrict__ send, unsigned long long sendCap,
	const unsigned* epoch)
{
	const int lane = threadIdx.x & 31;
	const int bank = bankBegin + blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
	if (bank >= bankEnd || bank * 32 >= count) return;
	const int local = bank * 32 + lane;
	float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
	if (local < count) rv = coarseR[begin - nVC + local];
	float4* second = nullptr;
	if (send) second = send + (unsigned long long)((*reinterpret_cast<const volatile unsigned*>(epoch) + 1u) & 1u) * sendCap;
	restrict_bank(goingNext, begin, count, bank, nVC, coarseR, lane, rv, second);
}

// The top levels hold a few hundred nodes: one CTA walks the remaining restrictions level by level, starting at
// firstLevel (level 2, or level 3 when level 2 is large enough to deserve the multi-CTA kernel above).
struct TopArgs
{
	int numLevel, nVC, firstLevel;
	int count[kMaxLevel + 1], begin[kMaxLevel + 1];
};
constexpr int kTopThreads = 256;    // small enough to be resident beside the persistent level-0 kernel (registers)
__global__ void __launch_bounds__(kTopThreads) restrict_top_kernel(const int* __restrict__ goingNext, TopArgs a, float4* __restrict__ coarseR)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = kTopThreads / 32;
	for (int level = a.firstLevel; level + 1 < a.numLevel; ++level)
	{
		const int banks = (a.count[level] + 31) >> 5;
		for (int bank = warp; bank < banks; bank += nWarps)
		{
			const int local = bank * 32 + lane;
			float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
			if (local < a.count[level]) rv = coarseR[a.begin[level] - a.nVC + local];
			restrict_bank(goingNext, a.begin[level], a.count[level], bank, a.nVC, coarseR, lane, rv);
		}
		__threadfence_block();
		__syncthreads();
	}
}

// One quarter of a packed block (the same four register batches solve_domain_ldg walks through one after the other):
// Q = 0,1,2: cyclic block diagonals 4Q+1 .. 4Q+4;  Q = 3: diagonals 13..15, the diagonal blocks and the half diagonal.
template <int Q>
__device__ __forceinline__ Vec3 solve_quarter(const float* __restrict__ blk, int lane, const Vec3 x)
{
	const float4* p4 = reinterpret_cast<const float4*>(blk) + lane;
	Vec3 y = { 0.f, 0.f, 0.f };
	float4 buf[9];
	if (Q < 3)
	{
#pragma unroll
		for (int q = 0; q < 9; ++q) buf[q] = ldg_stream4(p4 + 32 * (9 * Q + q));
		const float* m = reinterpret_cast<const float*>(buf);
#pragma unroll
		for (int dd = 0; dd < 4; ++dd) apply_block(m + 9 * dd, 4 * Q + 1 + dd, lane, x, y);
		return y;
	}
#pragma unroll
	for (int q = 0; q < 8; ++q) buf[q] = ldg_stream4(p4 + 32 * (27 + q));
	const float dTail = ldg_stream1(blk + kTailBase + lane);
	float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
	float h8 = 0.f;
	if (lane < 16)
	{
		const float4* ph = reinterpret_cast<const float4*>(blk + kHalfBase) + lane;
		h0 = ldg_stream4(ph);
		h1 = ldg_stream4(ph + 16);
		h8 = ldg_stream1(blk + kHalfTail + lane);
	}
	const float* m = reinterpret_cast<const float*>(buf);
#pragma unroll
	for (int dd = 0; dd < 3; ++dd) apply_block(m + 9 * dd, 13 + dd, lane, x, y);
	const float d0 = m[27], d1 = m[28], d2 = m[29], d3 = m[30], d4 = m[31], d5 = dTail;
	y.x = fmaf(d0, x.x, fmaf(d1, x.y, fmaf(d3, x.z, y.x)));
	y.y = fmaf(d1, x.x, fmaf(d2, x.y, fmaf(d4, x.z, y.y)));
	y.z = fmaf(d3, x.x, fmaf(d4, x.y, fmaf(d5, x.z, y.z)));
	const float b[9] = { h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h8 };
	const int peer = lane ^ 16;
	const float xjx = __shfl_sync(kFull, x.x, peer), xjy = __shfl_sync(kFull, x.y, peer), xjz = __shfl_sync(kFull, x.z, peer);
	float tx = 0.f, ty = 0.f, tz = 0.f;
	if (lane < 16)
	{
		y.x = fmaf(b[0], xjx, fmaf(b[1], xjy, fmaf(b[2], xjz, y.x)));
		y.y = fmaf(b[3], xjx, fmaf(b[4], xjy, fmaf(b[5], xjz, y.y)));
		y.z = fmaf(b[6], xjx, fmaf(b[7], xjy, fmaf(b[8], xjz, y.z)));
		tx = fmaf(b[0], x.x, fmaf(b[3], x.y, b[6] * x.z));
		ty = fmaf(b[1], x.x, fmaf(b[4], x.y, b[7] * x.z));
		tz = fmaf(b[2], x.x, fmaf(b[5], x.y, b[8] * x.z));
	}
	const float rx = __shfl_sync(kFull, tx, peer), ry = __shfl_sync(kFull, ty, peer), rz = __shfl_sync(kFull, tz, peer);
	if (lane >= 16) { y.x += rx; y.y += ry; y.z += rz; }
	return y;
}

// SchwarzLocalXSym on the blocks of levels >= 1 (cpp:1600-1696).  There are few of them (3 % of all blocks), so the kernel
// is latency-bound: FOUR warps share a block, each streaming one quarter of it in a single batch of loads, and the four
// partial products are added in a fixed order.
// The grid covers the coarse blocks this rank solves: `ownL1` level-1 blocks starting at l1Begin, then all blocks from
// topBegin on (levels >= 2); a single-GPU context owns every level-1 block.
// one coarse block by a group of four warps (w4 = warp within the group); `part`: the group's [3][32][3] scratch.  The caller
// puts a block-wide barrier between the two halves.
__device__ __forceinline__ Vec3 solve_coarse_part(const float* __restrict__ packed, const float4* __restrict__ coarseR, int blk, int lane, int w4,
	float (*part)[32][3])
{
	const float4 rv = coarseR[blk * 32 + lane];
	const Vec3 x = { rv.x, rv.y, rv.z };
	const float* base = packed + (size_t)blk * kTri;
	Vec3 y;
	if (w4 == 0) y = solve_quarter<0>(base, lane, x);
	else if (w4 == 1) y = solve_quarter<1>(base, lane, x);
	else if (w4 == 2) y = solve_quarter<2>(base, lane, x);
	else y = solve_quarter<3>(base, lane, x);
	if (w4 > 0) { part[w4 - 1][lane][0] = y.x; part[w4 - 1][lane][1] = y.y; part[w4 - 1][lane][2] = y.z; }
	return y;
}
__device__ __forceinline__ void solve_coarse_sum(Vec3 y, float4* __restrict__ coarseZ, int blk, int lane, int w4, float (*part)[32][3])
{
	if (w4 == 0)
	{
#pragma unroll
		for (int w = 0; w < 3; ++w) { y.x += part[w][lane][0]; y.y += part[w][lane][1]; y.z += part[w][lane][2]; }
		coarseZ[blk * 32 + lane] = make_float4(y.x, y.y, y.z, 0.f);
	}
}

__global__ void __launch_bounds__(128) solve_coarse_kernel(const float* __restrict__ packed, const float4* __restrict__ coarseR,
	float4* __restrict__ coarseZ, int l1Begin, int ownL1, int topBegin)
{
	__shared__ float part[3][32][3];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int blk = (int)blockIdx.x < ownL1 ? l1Begin + blockIdx.x : topBegin + (blockIdx.x - ownL1);
	const Vec3 y = solve_coarse_part(packed, coarseR, blk, lane, warp, part);
	__syncthreads();
	solve_coarse_sum(y, coarseZ, blk, lane, warp, part);
}

// Hierarchies with at most kTopSolveBlocks coarse blocks in all (meshes up to ~8k vertices, single GPU): the restrictions above
// level 1 and every coarse solve in ONE CTA - restrict_top_kernel's loop on 32 warps, a block barrier, then four warps per
// coarse block as in solve_coarse_kernel.  One launch less on a chain that is nothing but launch latency at this size (64x64
// cloth on a B200: 3.6 us against 2 x ~2 us, 10.6 -> 10.4 us per apply); the arithmetic and its order are those of the two
// kernels it replaces.
constexpr int kTopSolveBlocks = 8;
constexpr int kTopSolveThreads = kTopSolveBlocks * 128;
__global__ void __launch_bounds__(kTopSolveThreads) top_solve_kernel(const int* __restrict__ goingNext, TopArgs a, float4* __restrict__ coarseR,
	const float* __restrict__ packed, float4* __restrict__ coarseZ, int nBlocks)
{
	__shared__ float part[kTopSolveBlocks][3][32][3];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = kTopSolveThreads / 32;
	for (int level = a.firstLevel; level + 1 < a.numLevel; ++level)
	{
		const int banks = (a.count[level] + 31) >> 5;
		for (int bank = warp; bank < banks; bank += nWarps)
		{
			const int local = bank * 32 + lane;
			float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
			if (local < a.count[level]) rv = coarseR[a.begin[level] - a.nVC + local];
			restrict_bank(goingNext, a.begin[level], a.count[level], bank, a.nVC, coarseR, lane, rv);
		}
		__threadfence_block();
		__syncthreads();
	}
	const int grp = warp >> 2, w4 = warp & 3;
	Vec3 y = { 0.f, 0.f, 0.f };
	if (grp < nBlocks) y = solve_coarse_part(packed, coarseR, grp, lane, w4, part[grp]);
	__syncthreads();
	if (grp < nBlocks) solve_coarse_sum(y, coarseZ, grp, lane, w4, part[grp]);
}

// what CollectFinalZ (cpp:1698-1719) adds to every vertex below a level-1 node: Z_1 + Z_2[parent] + ...
__global__ void prolong_sum_kernel(const float4* __restrict__ coarseZ, const int* __restrict__ goingNext, int begin1, int first,
	int last, int nVC, int extraLevels, float4* __restrict__ zsum)
{
	const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= last) return;
	int node = begin1 + i;
	float4 z = coarseZ[node - nVC];
	for (int l = 0; l < extraLevels; ++l)
	{
		node = goingNext[node];
		const float4 u = coarseZ[node - nVC];
		z.x += u.x; z.y += u.y; z.z += u.z;
	}
	zsum[i] = make_float4(z.x, z.y, z.z, 0.f);
}

// SchwarzLocalXSym on level 0 fused with the level-0 gather of BuildResidualHierarchy and with CollectFinalZ.
// addCoarse = 0: the coarse levels are not ready (or absent); z holds the level-0 part only.
// One warp per bank.
__device__ __forceinline__ void solve_fine_bank(const float* __restrict__ packed, const float4* __restrict__ r, const int* __restrict__ s2o,
	const int* __restrict__ goingNext, const float4* __restrict__ zsum, int nv, int nVC, int bank, int packedBankBase, int addCoarse,
	float4* __restrict__ z, const int lane)
{
	const int v = bank * 32 + lane;
	const bool live = v < nv;
	int ov = 0, parent = 0;
	Vec3 x = { 0.f, 0.f, 0.f };
	if (live)
	{
		ov = s2o[v];
		const float4 rv = r[ov];
		x.x = rv.x; x.y = rv.y; x.z = rv.z;
		if (addCoarse) parent = goingNext[v];
	}
	Vec3 y = solve_domain_ldg(packed + (size_t)(bank - packedBankBase) * kTri, lane, x);
	if (live)
	{
		if (addCoarse)
		{
			const float4 c = zsum[parent - nVC];
			y.x += c.x; y.y += c.y; y.z += c.z;
		}
		z[ov] = make_float4(y.x, y.y, y.z, 0.f);
	}
}

__global__ void __launch_bounds__(kApplyThreads, 8) solve_fine_kernel(const float* __restrict__ packed,
	const float4* __restrict__ r, const int* __restrict__ s2o, const int* __restrict__ goingNext,
	const float4* __restrict__ zsum, int nv, int nVC, int bankBegin, int bankEnd, int packedBankBase, int addCoarse, float4* __restrict__ z)
{
	const int lane = threadIdx.x & 31;
	const int bank = bankBegin + blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
	if (bank >= bankEnd) return;
	solve_fine_bank(packed, r, s2o, goingNext, zsum, nv, nVC, bank, packedBankBase, addCoarse, z, lane);
}

// z += prolonged coarse solutions for the fine banks solved with addCoarse = 0
__global__ void add_coarse_kernel(const int* __restrict__ s2o, const int* __restrict__ goingNext, const float4* __restrict__ zsum,
	int vBegin, int vEnd, int nVC, float4* __restrict__ z)
{
	const int v = vBegin + blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= vEnd) return;
	const int ov = s2o[v];
	const float4 c = zsum[goingNext[v] - nVC];
	float4 y = z[ov];
	y.x += c.x; y.y += c.y; y.z += c.z;
	z[ov] = make_float4(y.x, y.y, y.z, 0.f);
}

// The same addition without prolong_sum — every vertex walks its own ancestors
// (CollectFinalZ as the reference writes it, cpp:1698-1719, through the Int4 ancestor table) and adds Z_1 + Z_2 + ... in
// prolong_sum's order, so the result is bit-identical while one launch leaves the latency-bound chain.  Used when the whole
// level-0 solve fits in the head of the apply graph (small meshes), where this kernel is all that follows the chain.
__global__ void add_coarse_walk_kernel(const int* __restrict__ s2o, const int4* __restrict__ coarseTables, const float4* __restrict__ coarseZ,
	int vBegin, int vEnd, int nVC, int levels, float4* __restrict__ z)
{
	const int v = vBegin + blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= vEnd) return;
	const int ov = s2o[v];
	const int4 a = coarseTables[v];
	float4 c = coarseZ[a.x - nVC];
	if (levels > 1) { const float4 u = coarseZ[a.y - nVC]; c.x += u.x; c.y += u.y; c.z += u.z; }
	if (levels > 2) { const float4 u = coarseZ[a.z - nVC]; c.x += u.x; c.y += u.y; c.z += u.z; }
	if (levels > 3) { const float4 u = coarseZ[a.w - nVC]; c.x += u.x; c.y += u.y; c.z += u.z; }
	float4 y = z[ov];
	y.x += c.x; y.y += c.y; y.z += c.z;
	z[ov] = make_float4(y.x, y.y, y.z, 0.f);
}

}  // namespace

#ifndef MAS_CPU_EMULATION   // host side: launches (the emulation has its own launcher)
// number of levels CollectFinalZ prolongs: l = 1 .. min(numLevel,4)-1 (cpp:1710, Q4) unless the fix is requested
static int prolonged_top(const Context* c)
{
	return c->optProlongAll ? c->numLevel : (c->numLevel < 4 ? c->numLevel : 4);
}

// Sharded context whose cuts are aligned to level-1 banks: every rank restricts its own level-1 banks to level 2 and the
// exchange carries level-2 residuals.  (With two levels there is nothing to exchange at all.)
static bool exchange_level2(const Context* c) { return c->world > 1 && c->alignedCuts; }

static PeerArgs peer_args(const Context* c)
{
	PeerArgs pa;
	for (int q = 0; q < kMaxWorld; ++q) { pa.send[q] = nullptr; pa.flags[q] = nullptr; }
	for (int q = 0; q <= kMaxWorld; ++q) pa.sliceBegin[q] = 0;
	for (int q = 0; q < c->world; ++q)
	{
		unsigned char* base = (unsigned char*)c->peerArena[q];
		pa.send[q] = reinterpret_cast<float4*>(base);
		pa.flags[q] = reinterpret_cast<unsigned*>(base + 2 * sizeof(float4) * c->arenaCap);
	}
	// what the peers pull: level-1 residuals (cuts not aligned) or level-2 residuals (aligned cuts), as coarse indices
	const bool l2x = exchange_level2(c);
	for (int q = 0; q <= c->world; ++q)
		pa.sliceBegin[q] = l2x ? c->levelSize[2][1] - c->nVC + c->l2Slice[q] : c->levelSize[1][1] - c->nVC + c->l1Slice[q];
	unsigned* ctl = reinterpret_cast<unsigned*>((unsigned char*)c->peerArena[c->rank] + 2 * sizeof(float4) * c->arenaCap);
	pa.epoch = ctl + kMaxWorld;
	pa.ticket = ctl + kMaxWorld + 1;
	pa.error = c->peerErrDev;      // page-locked host word: the host sees a timed-out wait without a device round trip
	pa.cap = c->arenaCap;
	pa.world = c->world;
	pa.rank = c->rank;
	pa.strictPublish = c->optStrictPublish;
	return pa;
}

int apply_begin(Context* c, const float4* r)
{
	cudaStream_t st = c->stream;
	const int ownBanks = c->ownFineEnd - c->ownFineBegin;
	if (c->numLevel < 2) return MAS_OK;
	const bool peers = use_peers(c);
	if (c->world > 1 && !peers)
	{
		MAS_CUDA(c, cudaMemsetAsync(c->coarseR.p, 0, sizeof(float4) * (size_t)c->nCoarseNodes, st));
	}
	const bool l2x = exchange_level2(c);
	PeerArgs pa;
	if (peers) pa = peer_args(c);
	if (ownBanks > 0)
	{
		const bool publish = peers && !l2x;
		restrict_fine_kernel<<<cdiv(ownBanks, kWarpsPerCta * kRestrictBanks), kApplyThreads, 0, st>>>(r, c->s2o.p, c->goingNext.p, c->nv, c->nVC,
			c->ownFineBegin, c->ownFineEnd, c->coarseR.p, publish ? pa.send[c->rank] : nullptr, publish ? pa.cap : 0ull,
			publish ? pa.epoch : nullptr);
		c->applyLaunches += 1;
	}
	if (l2x && c->numLevel > 2 && c->l1BlockEnd > c->l1BlockBegin)
	{
		// own level-1 banks -> level 2 (complete: no level-1 bank straddles a cut); published for the peers
		const int cnt1 = c->levelSize[1][0], begin1 = c->levelSize[1][1];
		restrict_l1_kernel<<<cdiv(c->l1BlockEnd - c->l1BlockBegin, kWarpsPerCta), kApplyThreads, 0, st>>>(c->goingNext.p, begin1, cnt1, c->nVC,
			c->l1BlockBegin, c->l1BlockEnd, c->coarseR.p, peers ? pa.send[c->rank] : nullptr, peers ? pa.cap : 0ull,
			peers ? pa.epoch : nullptr);
		c->applyLaunches += 1;
	}
	return MAS_OK;
}

// coarse levels: needs the complete level-1 residuals in coarseR (after the exchange when world > 1)
// skipProlongSum: the caller adds the coarse part with add_coarse_walk (which walks the ancestors itself)
static int launch_coarse(Context* c, cudaStream_t st, bool skipProlongSum = false)
{
	if (c->numLevel < 2) return MAS_OK;
	const int cnt1 = c->levelSize[1][0], begin1 = c->levelSize[1][1];
	const int nCoarseBlocks = c->nCoarseNodes / 32;
	const bool l2x = exchange_level2(c);
	if (use_peers(c) && !(l2x && c->numLevel < 3))
	{
		const int first = l2x ? c->levelSize[2][1] - c->nVC : begin1 - c->nVC;
		const int count = l2x ? c->levelSize[2][0] : cnt1;
		int grid = cdiv(count, 256 * kGatherPerThread);
		if (grid > 128) grid = 128;
		if (grid < 1) grid = 1;
		gather_peers_kernel<<<grid, 256, 0, st>>>(peer_args(c), first, count, c->coarseR.p);
		c->applyLaunches += 1;
	}
	// small single-GPU meshes (at most 16 level-1 banks): the one-CTA kernel that walks the top levels starts at level 1
	// already, which takes restrict_l1 — one launch — off the latency-bound chain (bit-identical, measured on a B200:
	// together with the ancestor walk below 12.5 -> 10.5 us per apply on the 64x64 cloth)
	const bool topFromL1 = c->world == 1 && c->numLevel > 2 && cnt1 <= 512;
	bool fusedSolve = false;
	if (c->numLevel > 2 && !l2x && !topFromL1)
	{
		restrict_l1_kernel<<<cdiv(cdiv(cnt1, 32), kWarpsPerCta), kApplyThreads, 0, st>>>(c->goingNext.p, begin1, cnt1, c->nVC, 0,
			cdiv(cnt1, 32), c->coarseR.p, nullptr, 0ull, nullptr);
		c->applyLaunches += 1;
	}
	if (c->numLevel > 3 || topFromL1)
	{
		TopArgs a;
		a.numLevel = c->numLevel; a.nVC = c->nVC; a.firstLevel = topFromL1 ? 1 : 2;
		for (int l = 0; l <= kMaxLevel; ++l) { a.count[l] = 0; a.begin[l] = 0; }
		for (int l = 1; l <= c->numLevel; ++l) { a.count[l] = c->levelSize[l][0]; a.begin[l] = c->levelSize[l][1]; }
		const int cnt2 = c->levelSize[2][0];
		if (cnt2 > 2048 && !topFromL1)
		{
			// a large level 2 (meshes beyond ~2M vertices): level 2 -> 3 on many CTAs, the single CTA takes over from level 3
			restrict_l1_kernel<<<cdiv(cdiv(cnt2, 32), kWarpsPerCta), kApplyThreads, 0, st>>>(c->goingNext.p, c->levelSize[2][1], cnt2, c->nVC, 0,
				cdiv(cnt2, 32), c->coarseR.p, nullptr, 0ull, nullptr);
			c->applyLaunches += 1;
			a.firstLevel = 3;
		}
		if (topFromL1 && nCoarseBlocks <= kTopSolveBlocks)
		{
			top_solve_kernel<<<1, kTopSolveThreads, 0, st>>>(c->goingNext.p, a, c->coarseR.p,
				c->packedInv.p + (size_t)(c->ownFineEnd - c->ownFineBegin) * kTri, c->coarseZ.p, nCoarseBlocks);
			c->applyLaunches += 1;
			fusedSolve = true;
		}
		else if (a.firstLevel + 1 < c->numLevel)
		{
			restrict_top_kernel<<<1, kTopThreads, 0, st>>>(c->goingNext.p, a, c->coarseR.p);
			c->applyLaunches += 1;
		}
	}
	// level-1 blocks stay partitioned (each rank solves the blocks that hold its own level-1 nodes); levels >= 2 are solved
	// redundantly on every rank, which removes any exchange of z
	const int ownL1 = c->l1BlockEnd - c->l1BlockBegin;
	const int solved = ownL1 + (nCoarseBlocks - c->nL1Blocks);
	if (solved > 0 && !fusedSolve)
	{
		solve_coarse_kernel<<<solved, 128, 0, st>>>(c->packedInv.p + (size_t)(c->ownFineEnd - c->ownFineBegin) * kTri, c->coarseR.p,
			c->coarseZ.p, c->l1BlockBegin, ownL1, c->nL1Blocks);
		c->applyLaunches += 1;
	}
	const int first = c->world > 1 ? c->l1Slice[c->rank] : 0, last = c->world > 1 ? c->l1Slice[c->rank + 1] : cnt1;
	if (last > first && !skipProlongSum)
	{
		prolong_sum_kernel<<<cdiv(last - first, 256), 256, 0, st>>>(c->coarseZ.p, c->goingNext.p, begin1, first, last, c->nVC,
			prolonged_top(c) - 2, c->coarseZsum.p);
		c->applyLaunches += 1;
	}
	return MAS_OK;
}

static void launch_fine(Context* c, cudaStream_t st, const float4* r, float4* z, int bankBegin, int bankEnd, int addCoarse)
{
	if (bankEnd <= bankBegin) return;
	solve_fine_kernel<<<cdiv(bankEnd - bankBegin, kWarpsPerCta), kApplyThreads, 0, st>>>(c->packedInv.p, r, c->s2o.p, c->goingNext.p,
		c->coarseZsum.p, c->nv, c->nVC, bankBegin, bankEnd, c->ownFineBegin, addCoarse, z);
	c->applyLaunches += 1;
}

int apply_end(Context* c, const float4* r, float4* z)
{
	cudaStream_t st = c->stream;
	if (int rc = launch_coarse(c, st)) return rc;
	if (c->ownFineEnd > c->ownFineBegin)
	{
		if (c->optTimeKernels) MAS_CUDA(c, cudaEventRecord(c->evF0, st));
		launch_fine(c, st, r, z, c->ownFineBegin, c->ownFineEnd, prolonged_top(c) >= 2 ? 1 : 0);
		if (c->optTimeKernels) MAS_CUDA(c, cudaEventRecord(c->evF1, st));
	}
	MAS_CUDA(c, cudaGetLastError());
	return MAS_OK;
}

// Whole apply as a two-branch capture (graph path; single GPU, or sharded with the peer-memory exchange).  `st` is the capturing origin stream; side streams and
// events belong to the context.  headBanks fine banks are solved concurrently with the coarse chain.
int apply_forked(Context* c, const float4* r, float4* z, cudaStream_t st)
{
	const int top = prolonged_top(c);
	const int ownBanks = c->ownFineEnd - c->ownFineBegin;
	int head = 0, headChunk = 1 << 30;
	if (top >= 2 && c->optApplyVariant != 0)
	{
		if (c->optApplyVariant > 0)
			head = (int)((long long)ownBanks * c->optApplyVariant / 1000);
		else
		{
			// auto: as many banks as stream in the time the coarse chain takes.  The chain costs a fixed latency (five
			// dependent launches, the peer wait), the level-0 restriction of the owned vertices and the coarse solves above them
			// (level 1 is partitioned like the fine banks), all of it slowed by the streaming kernel it shares HBM with; ~330 banks
			// stream per microsecond.  Fitted on sweeps at 1M vertices per GPU (optimum ~45 % of the banks on 1 and 2 GPUs:
			// profiles/r02_head_sweep.txt; the round-1 fit, 2400 + 0.004 ownVerts, was made under the one-wave limit).
			const long long ownVerts = 32ll * ownBanks;
			head = (int)(2400 + 23 * ownVerts / 2000);
			if (head > 15000) head = 15000;   // beyond ~1M vertices the chain grows little while every extra one-wave kernel costs a bubble (4.2M: 427 us at 15 %, 445 us at 38 %)
			if (use_peers(c)) head += 1500;   // signal + peer wait + pull over NVLink
		}
		// The CTA dispatcher works through grids in launch order, and stream / node priority does not let a later grid overtake
		// the not-yet-dispatched CTAs of an earlier one (measured: beside a head of more than one wave, 148 x 8 CTAs, every
		// chain kernel waited 20-26 us until the head's last CTA had been dispatched).  The head therefore never exceeds one
		// wave minus room for the chain's own CTAs: it is resident in full right away and the chain is dispatched beside it.
		// (Round 2 also measured the whole chain as ONE persistent kernel with grid barriers, launched ahead of the head: it
		// removes the dispatch queueing and the launch gaps, but with at most one CTA per SM its restriction and solve phases
		// lose their parallelism and its resident CTAs take registers from the streaming kernel: 117.0 against 104.5 us at 1M
		// vertices, 37.2 / 30.5 us at 262k, 12.4 / 10.4 us at 4k; profiles/r02_fused_coarse_chain_kernel_measurement.txt.  Removed.)
		// A head that has to cover a longer chain (sharded contexts: the peer exchange is on it) is launched as SEVERAL one-wave
		// kernels one after the other on the side stream: each is dispatched in full the moment its predecessor has drained, so
		// a chain kernel never finds undispatched head CTAs ahead of it.
		headChunk = (c->smCount * 8 - 80) * kWarpsPerCta;
		head = (head + kWarpsPerCta - 1) / kWarpsPerCta * kWarpsPerCta;
		if (head > ownBanks) head = ownBanks;
	}
	if (head == 0)
	{
		cudaStream_t saved = c->stream;
		c->stream = st;
		int rc = apply_begin(c, r);
		if (rc == MAS_OK) rc = apply_end(c, r, z);
		c->stream = saved;
		return rc;
	}
	const int b0 = c->ownFineBegin, b1 = b0 + head, b2 = c->ownFineEnd;
	// small meshes: the whole level-0 solve is in the head, so add_coarse is all that follows the chain and walks the
	// ancestors itself (no prolong_sum: one launch less on the chain)
	const bool walk = c->world == 1 && head == ownBanks;
	MAS_CUDA(c, cudaEventRecord(c->evFork, st));
	MAS_CUDA(c, cudaStreamWaitEvent(c->sideA, c->evFork, 0));
	for (int hb = b0; hb < b1; hb += headChunk)                      // level-0 part only, no coarse data needed
		launch_fine(c, c->sideA, r, z, hb, hb + headChunk < b1 ? hb + headChunk : b1, 0);
	MAS_CUDA(c, cudaEventRecord(c->evHead, c->sideA));
	{
		cudaStream_t saved = c->stream;
		c->stream = st;
		int rc = apply_begin(c, r);
		if (rc == MAS_OK) rc = launch_coarse(c, st, walk);
		c->stream = saved;
		if (rc != MAS_OK) return rc;
	}
	MAS_CUDA(c, cudaEventRecord(c->evCoarse, st));
	MAS_CUDA(c, cudaStreamWaitEvent(c->sideB, c->evCoarse, 0));
	launch_fine(c, c->sideB, r, z, b1, b2, 1);
	MAS_CUDA(c, cudaEventRecord(c->evTail, c->sideB));
	MAS_CUDA(c, cudaStreamWaitEvent(st, c->evHead, 0));
	{
		const int vBegin = b0 * 32, vEnd = b1 * 32 < c->nv ? b1 * 32 : c->nv;
		if (vEnd > vBegin)
		{
			if (walk)
				add_coarse_walk_kernel<<<cdiv(vEnd - vBegin, 256), 256, 0, st>>>(c->s2o.p, c->coarseTables.p, c->coarseZ.p, vBegin, vEnd, c->nVC,
					top - 1, z);
			else
				add_coarse_kernel<<<cdiv(vEnd - vBegin, 256), 256, 0, st>>>(c->s2o.p, c->goingNext.p, c->coarseZsum.p, vBegin, vEnd, c->nVC, z);
			c->applyLaunches += 1;
		}
	}
	MAS_CUDA(c, cudaStreamWaitEvent(st, c->evTail, 0));
	MAS_CUDA(c, cudaGetLastError());
	return MAS_OK;
}

// Kernel nodes of the captured apply graph: the latency-bound coarse chain gets the highest launch priority, the two
// streaming kernels (level-0 solve, coarse addition) the lowest, so that a chain kernel's few CTAs are dispatched as soon as
// any CTA slot frees up instead of queueing behind the not-yet-dispatched CTAs of the concurrent level-0 solve.
int prioritize_apply_graph(Context* c, cudaGraph_t graph)
{
	int prLow = 0, prHigh = 0;
	MAS_CUDA(c, cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
	size_t n = 0;
	MAS_CUDA(c, cudaGraphGetNodes(graph, nullptr, &n));
	std::vector<cudaGraphNode_t> nodes(n);
	if (n) MAS_CUDA(c, cudaGraphGetNodes(graph, nodes.data(), &n));
	for (size_t i = 0; i < n; ++i)
	{
		cudaGraphNodeType type;
		MAS_CUDA(c, cudaGraphNodeGetType(nodes[i], &type));
		if (type != cudaGraphNodeTypeKernel) continue;
		cudaKernelNodeParams kp;
		MAS_CUDA(c, cudaGraphKernelNodeGetParams(nodes[i], &kp));
		const bool streaming = kp.func == (void*)solve_fine_kernel || kp.func == (void*)add_coarse_kernel ||
			kp.func == (void*)add_coarse_walk_kernel;
		cudaKernelNodeAttrValue v;
		v.priority = streaming ? prLow : prHigh;
		MAS_CUDA(c, cudaGraphKernelNodeSetAttribute(nodes[i], cudaKernelNodeAttributePriority, &v));
	}
	return MAS_OK;
}

#endif  // MAS_CPU_EMULATION

}  // namespace mas
