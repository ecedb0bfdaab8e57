// Device-resident preconditioned conjugate gradients: the immediate CALLER of Preconditioning().
// The reference ships no solver (its caller owns the PCG loop, SeSchwarzPreconditioner.h:55-63); BASELINE config 2
// ("PCG to 1e-5 residual with MAS, iteration count and wall time") needs one on both sides, and without it the
// host<->device copies of r and z would dwarf the apply.
//
//   A is the caller's Hessian in the same arrays PreparePreconditioner receives (original vertex order):
//   diag[nv] + offdiag[nnz] column-major 3x3 blocks, CSR adjacency (ranges, idx); vectors are 16-byte xyzw.
//   x0 = 0;  stop when ||r||_2 / ||b||_2 < relTol;  dot products in FP64 with a fixed reduction order.
//
// One iteration = 4 launches + the apply's own launches, captured once in a CUDA graph and replayed; the stopping test is
// evaluated on the device (a flag turns the remaining launches of a batch into no-ops), so the host synchronises once per
// batch of iterations, not once per iteration.  Every kernel is launched with exactly as many CTAs as are resident at once
// (occupancy x SM count) and strides over its rows: with one CTA per 256 rows the 1M-vertex kernels ran 1.4-2.3 waves and the
// last, partly filled wave cost 20 % of each (profiles/r02_pcg_iteration_timeline.txt).
//   A includes the collision Hessians of the stencils of the last PreparePreconditioner (what the preconditioner was built
//   for, cpp:1164-1227): stiff (w (x) w) (x) (d d^T) per stencil, applied matrix-free by stencil_spmv (one thread per stencil,
//   float atomics into Ap: with stencils the summation order, and with it the last bit of the iterates, varies from run to run).
//   spmv_dot      Ap = A p, partial sums of p.Ap.  A is converted once per solve to a sliced-ELL layout (32-row slices,
//                 every (block slot, entry) of a slice is 32 consecutive floats; slot 0 is the diagonal block): lane = row,
//                 all loads coalesced, nine FMAs per eleven loads and no cross-lane traffic (the CSR kernel it replaces
//                 staged blocks through shared memory and needed a segmented shuffle scan per 32 blocks: 112 us vs 397 MB
//                 at 1M vertices)
//   axpy_rr       alpha = rz / p.Ap;  r -= alpha Ap;  partial sums of r.r
//   (apply)       z = M^-1 r
//   dot_rz        partial sums of r.z
//   update_p      x += alpha p;  beta = rz' / rz;  p = z + beta p  (p is read once for both); the last CTA evaluates the
//                 stopping test for this iteration
//   The vector kernels issue their first loads BEFORE they reduce the previous pass's partial sums, so the reduction's
//   latency (dependent L2 reads, two barriers) is hidden under the memory round trip instead of preceding it.
#include "mas_internal.h"

namespace mas {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kPcgThreads = 256;
constexpr int kPcgWarps = kPcgThreads / 32;
constexpr int kMaxPartials = 2048;  // upper bound of CTAs per reduction pass (the grids are occupancy x SM count)
// SpMV staging (spmv_dot_kernel): every warp streams its part of the ELL image through its own ring of shared-memory
// buffers, filled by bulk asynchronous copies (TMA) that complete on an mbarrier
constexpr int kSpmvThreads = 512;
constexpr int kSpmvWarps = kSpmvThreads / 32;
constexpr int kSlotWords = 320;      // one block slot of a slice: 9 x 32 values + 32 column indices
constexpr int kChunkSlots = 3;       // slots per bulk copy (3,840 bytes)
constexpr int kSpmvStages = 3;       // buffers per warp: two copies in flight while the third is consumed
constexpr int kChunkWords = kChunkSlots * kSlotWords;
constexpr int kGatherAhead = 3;      // chunks between the gathers of p and their use
constexpr int kQRing = kGatherAhead + 1;
constexpr int kIdxRing = 2;          // chunks between the loads of the column indices and the gathers that use them
constexpr int kIdxAhead = kGatherAhead + kIdxRing;
constexpr int kSpmvUnroll = 4;       // a common multiple of both ring lengths
constexpr int kSpmvPrologue = 8;     // steps before chunk 0 (>= kIdxAhead, a multiple of kSpmvUnroll)
static_assert(kSpmvUnroll % kQRing == 0 && kSpmvUnroll % kIdxRing == 0 && kSpmvPrologue % kSpmvUnroll == 0 && kSpmvPrologue >= kIdxAhead, "");
constexpr size_t kSpmvSmemBytes = (size_t)kSpmvWarps * kSpmvStages * kChunkWords * sizeof(float);
constexpr int kVecUnroll = 4;       // elements per thread and trip of the vector kernels, loads in flight together

struct PcgState
{
	double rz, rr, rr0, alpha;   // alpha: step length of the running iteration (axpy_rr -> update_p)
	int done, iters, it;         // done: 1 = converged, 2 = maxIter reached
	int pad;                     // update_p's ticket counter
	int stageErr, pad2;          // spmv_dot: a bulk copy never completed
};

__device__ __forceinline__ double block_sum(double v, double* sh)
{
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	__syncthreads();
	if (lane == 0) sh[warp] = v;
	__syncthreads();
	double t = 0.0;
	for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];  // every thread, same order
	return t;
}

// every CTA reduces the (<= kMaxPartials) partial sums of the previous pass in the same fixed order
__device__ __forceinline__ double reduce_partials(const double* __restrict__ partials, int n, double* sh)
{
	double v = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
	return block_sum(v, sh);
}

// ---- sliced-ELL copy of the caller's block CSR (once per solve) ---------------------------------------------------------
// slice g = rows 32g .. 32g+31, width w_g = 1 + its longest row (in slots); slot k of slice g is one 1,280-byte record
//   ell[kSlotWords (sliceStart[g] + k) + 32 e + lane]   e = 0..8: entry e of the column-major 3x3 block of row 32g+lane
//                                                      e = 9   : its column vertex (int bits; -1 = padding, zero block)
// k = 0 is the row's diagonal block, k >= 1 its (k-1)-th off-diagonal block.  The slices follow each other without gaps, so
// any range of slots - across slice boundaries too - is one contiguous, 16-byte aligned piece of memory.
__global__ void ell_width_kernel(const int* __restrict__ ranges, int nv, int nSlices, int* __restrict__ sliceSlots)
{
	const int lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (g > nSlices) return;
	const int row = g * 32 + lane;
	int deg = row < nv ? ranges[row + 1] - ranges[row] : 0;
	for (int off = 16; off > 0; off >>= 1) deg = max(deg, __shfl_xor_sync(kFull, deg, off));
	if (lane == 0) sliceSlots[g] = g < nSlices ? deg + 1 : 0;   // entry nSlices: the scan leaves the total there
}

__global__ void ell_fill_kernel(const float* __restrict__ diag, const float* __restrict__ off, const int* __restrict__ ranges,
	const int* __restrict__ idx, int nv, const int* __restrict__ sliceStart, float* __restrict__ ell)
{
	const int lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (g * 32 >= nv) return;
	const int row = g * 32 + lane;
	const int rs = row < nv ? ranges[row] : 0, re = row < nv ? ranges[row + 1] : 0;
	const int base = sliceStart[g], width = sliceStart[g + 1] - base;
	for (int k = 0; k < width; ++k)
	{
		const bool has = k == 0 ? row < nv : rs + k - 1 < re;
		const float* m = k == 0 ? diag + 9 * (size_t)row : off + 9 * (size_t)(rs + k - 1);
		float* dst = ell + (size_t)kSlotWords * (base + k) + lane;
#pragma unroll
		for (int e = 0; e < 9; ++e) dst[32 * e] = has ? m[e] : 0.0f;
		dst[32 * 9] = __int_as_float(!has ? -1 : k == 0 ? row : idx[rs + k - 1]);
	}
}

// two arrays of partial sums at once (one pair of barriers instead of two)
__device__ __forceinline__ void reduce_partials2(const double* __restrict__ pa, const double* __restrict__ pb, int n, double* sh,
	double& ra, double& rb)
{
	double a = 0.0, b = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) { a += pa[i]; b += pb[i]; }
	for (int off = 16; off > 0; off >>= 1)
	{
		a += __shfl_xor_sync(kFull, a, off);
		b += __shfl_xor_sync(kFull, b, off);
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
	__syncthreads();
	if (lane == 0) { sh[warp] = a; sh[nw + warp] = b; }
	__syncthreads();
	ra = 0.0; rb = 0.0;
	for (int w = 0; w < nw; ++w) { ra += sh[w]; rb += sh[nw + w]; }  // every thread, same order
}

// ---- bulk asynchronous copies global -> shared (TMA, cp.async.bulk) completing on an mbarrier ----------------------------
#ifdef MAS_CPU_EMULATION
__device__ __forceinline__ void stage_init(unsigned long long*) {}
__device__ __forceinline__ void stage_copy(float* dst, const float* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ bool stage_wait(unsigned long long*, unsigned) { return true; }
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_init(unsigned long long* bar)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
}
// one thread: expect `bytes` on the barrier, then start the copy (src, dst and bytes are multiples of 16)
__device__ __forceinline__ void stage_copy(float* dst, const float* src, unsigned bytes, unsigned long long* bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
		"r"(bytes), "r"(smem_u32(bar))
		: "memory");
}
// bounded: a protocol bug must not hang the GPU
__device__ __forceinline__ bool stage_wait(unsigned long long* bar, unsigned parity)
{
	const unsigned a = smem_u32(bar);
#pragma unroll 1
	for (unsigned spin = 0; spin < (1u << 26); ++spin)
	{
		unsigned done;
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}\n"
			: "=r"(done)
			: "r"(a), "r"(parity)
			: "memory");
		if (done) return true;
	}
	return false;
}
#endif

#ifdef MAS_CPU_EMULATION
#define MAS_PCG_DYNAMIC_SMEM(name) unsigned char* name = emu_dynamic_smem()
#else
#define MAS_PCG_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

// Ap = A p and partial p.Ap.  The slices are dealt to the warps of the grid in contiguous runs; a warp's run is one contiguous
// stretch of the ELL image, which it pulls through its ring of kSpmvStages shared-memory buffers in chunks of kChunkSlots
// slots: lane 0 issues the bulk copies (two chunks ahead), all lanes wait on the chunk's mbarrier and read their own column
// of it (lane = row, conflict-free).  Nothing of A lands in registers before it is used, so the bytes in flight per SM are
// bounded by shared memory (2/3 of 180 KB), not by the register file: the register-staged kernel this replaces (40 loads per
// warp and batch, then the dependent gathers of p) held 5.4 TB/s at 50 % occupancy, stalled on the scoreboard and on the
// load/store queue (profiles/r02_spmv_register_staged_ncu.txt).
// Slice boundaries may fall inside a chunk.
__global__ void __launch_bounds__(kSpmvThreads, 1) spmv_dot_kernel(const int* __restrict__ sliceStart, const float* __restrict__ ell,
	const float4* __restrict__ p, float4* __restrict__ Ap, int nv, double* __restrict__ partials, const volatile PcgState* st,
	int* __restrict__ err)
{
	MAS_PCG_DYNAMIC_SMEM(smem);
	__shared__ unsigned long long bars[kSpmvWarps * kSpmvStages];
	__shared__ double sh[kSpmvWarps];
	if (st->done) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nSlices = (nv + 31) >> 5;
	const long long nWarps = (long long)gridDim.x * kSpmvWarps, wg = (long long)blockIdx.x * kSpmvWarps + warp;
	const int s0 = (int)(nSlices * wg / nWarps), s1 = (int)(nSlices * (wg + 1) / nWarps);
	float* ring = reinterpret_cast<float*>(smem) + (size_t)warp * kSpmvStages * kChunkWords;
	unsigned long long* bar = bars + warp * kSpmvStages;
	if (lane == 0)
		for (int i = 0; i < kSpmvStages; ++i) stage_init(bar + i);
#ifndef MAS_CPU_EMULATION
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
	__syncwarp();

	double dot = 0.0;
	if (s1 > s0)
	{
		const int slot0 = sliceStart[s0], nSlots = sliceStart[s1] - slot0;
		const int nChunks = (nSlots + kChunkSlots - 1) / kChunkSlots;
		const float* src = ell + (size_t)kSlotWords * slot0;
		auto issue = [&](int c) {   // lane 0
			const int n = nSlots - c * kChunkSlots < kChunkSlots ? nSlots - c * kChunkSlots : kChunkSlots;
			stage_copy(ring + (c % kSpmvStages) * kChunkWords, src + (size_t)c * kChunkWords, (unsigned)(n * kSlotWords * sizeof(float)),
				bar + c % kSpmvStages);
		};
		if (lane == 0)
			for (int c = 0; c < kSpmvStages && c < nChunks; ++c) issue(c);
		// widths of the run's slices, 32 at a time in the lanes
		int slice = s0, metaBase = s0;
		int meta = s0 + lane < s1 ? sliceStart[s0 + lane + 1] - sliceStart[s0 + lane] : 0;
		int width = __shfl_sync(kFull, meta, 0), kIn = 0;
		float yx = 0.f, yy = 0.f, yz = 0.f;
		float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
		bool ok = true;
		// Three software pipelines per warp, all in units of chunks: bulk copies (kSpmvStages - 1 ahead, into shared memory),
		// column indices (kIdxAhead ahead, plain loads of the same records' index words) and the gathers of p
		// (kGatherAhead ahead, into a register ring).  Chunk c is multiplied when its copy has landed and its gathers, issued
		// three steps earlier, have returned: no step waits on a load issued in the same step.  (With the gathers only one
		// chunk ahead the per-chunk step took one L2 round trip under load, ~1 us, and the kernel was latency-bound at 82 us.)
		int ci[kIdxRing][kChunkSlots];
		float4 q[kQRing][kChunkSlots];
		for (int c0 = -kSpmvPrologue; c0 < nChunks; c0 += kSpmvUnroll)
		{
#pragma unroll
			for (int j = 0; j < kSpmvUnroll; ++j)
			{
				const int c = c0 + j;   // c mod kSpmvUnroll == j: every ring index below is a compile-time constant
				const int cg = c + kGatherAhead, cx = c + kIdxAhead;
				if (cg >= 0 && cg < nChunks)
				{
#pragma unroll
					for (int u = 0; u < kChunkSlots; ++u)
					{
						const int col = ci[(j + kGatherAhead) % kIdxRing][u];
						q[(j + kGatherAhead) % kQRing][u] = p[col < 0 ? 0 : col];
					}
				}
				if (cx >= 0 && cx < nChunks)
				{
#pragma unroll
					for (int u = 0; u < kChunkSlots; ++u)
					{
						const int slot = cx * kChunkSlots + u;
						ci[(j + kIdxAhead) % kIdxRing][u] = slot < nSlots ? __float_as_int(src[(size_t)slot * kSlotWords + 9 * 32 + lane]) : -1;
					}
				}
				if (c < 0 || c >= nChunks) continue;
				ok = stage_wait(bar + c % kSpmvStages, (unsigned)(c / kSpmvStages) & 1u) && ok;
				__syncwarp();
				const float* buf = ring + (c % kSpmvStages) * kChunkWords;
				const int n = nSlots - c * kChunkSlots;
#pragma unroll
				for (int u = 0; u < kChunkSlots; ++u)
				{
					if (u >= n) break;
					const float* m = buf + u * kSlotWords + lane;
					const float4 pv = q[j % kQRing][u];
					if (kIn == 0) xv = pv;                       // slot 0 is the diagonal block: its column is the row itself
					yx = fmaf(m[0], pv.x, fmaf(m[96], pv.y, fmaf(m[192], pv.z, yx)));
					yy = fmaf(m[32], pv.x, fmaf(m[128], pv.y, fmaf(m[224], pv.z, yy)));
					yz = fmaf(m[64], pv.x, fmaf(m[160], pv.y, fmaf(m[256], pv.z, yz)));
					if (++kIn == width)
					{
						const int row = slice * 32 + lane;
						if (row < nv)
						{
							Ap[row] = make_float4(yx, yy, yz, 0.f);
							dot += (double)xv.x * yx + (double)xv.y * yy + (double)xv.z * yz;
						}
						yx = yy = yz = 0.f;
						kIn = 0;
						++slice;
						if (slice - metaBase == 32 && slice < s1)
						{
							metaBase = slice;
							meta = slice + lane < s1 ? sliceStart[slice + lane + 1] - sliceStart[slice + lane] : 0;
						}
						width = __shfl_sync(kFull, meta, (slice - metaBase) & 31);
					}
				}
				// the buffer is free: every lane has read its column of chunk c
				__syncwarp();
				if (lane == 0 && c + kSpmvStages < nChunks) issue(c + kSpmvStages);
			}
		}
		if (!ok && lane == 0) *err = 1;
	}
	const double t = block_sum(dot, sh);
	if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// Ap += (sum of the collision Hessians) p and the matching part of p.Ap: stencil s contributes stiff (w (x) w) (x) (d d^T)
// (PrepareCollisionHessian, cpp:1201-1227: OuterProduct(d, d stiff) weighted by w_a w_b for every vertex pair of the stencil),
// i.e. alpha = stiff sum_k w_k (d . p_k),  Ap_k += w_k alpha d,  p.Ap += alpha^2 / stiff.  Stencil indices are original ids.
__global__ void __launch_bounds__(kPcgThreads) stencil_spmv_kernel(const Stencil* __restrict__ stencils, int nStencil,
	const float4* __restrict__ p, float4* __restrict__ Ap, double* __restrict__ partials, const volatile PcgState* st)
{
	__shared__ double sh[kPcgWarps];
	if (st->done) return;
	double dot = 0.0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nStencil; i += gridDim.x * blockDim.x)
	{
		const Stencil s = stencils[i];
		float a = 0.f;
		for (int k = 0; k < s.n; ++k)
		{
			const float4 q = p[s.index[k]];
			a = fmaf(s.weight[k], fmaf(s.dir[0], q.x, fmaf(s.dir[1], q.y, s.dir[2] * q.z)), a);
		}
		const float alpha = s.stiff * a;
		for (int k = 0; k < s.n; ++k)
		{
			const float w = s.weight[k] * alpha;
			float* dst = reinterpret_cast<float*>(Ap + s.index[k]);
			atomicAdd(dst + 0, w * s.dir[0]);
			atomicAdd(dst + 1, w * s.dir[1]);
			atomicAdd(dst + 2, w * s.dir[2]);
		}
		dot += (double)alpha * (double)a;
	}
	const double t = block_sum(dot, sh);
	if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// r -= alpha Ap, partial sums of r.r; alpha = rz / p.Ap is left in the state for update_p (x += alpha p happens there,
// where p is read anyway)
__global__ void __launch_bounds__(kPcgThreads) axpy_rr_kernel(float4* __restrict__ r, const float4* __restrict__ Ap, int nv,
	const double* __restrict__ pApPartials, const double* __restrict__ pApStencilPartials, int nPartials,
	double* __restrict__ rrPartials, PcgState* stw)
{
	__shared__ double sh[kPcgWarps];
	const volatile PcgState* st = stw;
	if (st->done) return;
	const int stride = gridDim.x * blockDim.x;
	int i0 = blockIdx.x * blockDim.x + threadIdx.x;
	float4 rv[kVecUnroll], av[kVecUnroll];
	// all loads of the kVecUnroll elements are issued before the first use (one memory round trip, not four), and the first
	// trip's before the reduction below
#pragma unroll
	for (int u = 0; u < kVecUnroll; ++u)
	{
		const int i = i0 + u * stride;
		if (i < nv) { rv[u] = r[i]; av[u] = Ap[i]; }
	}
	double pAp = reduce_partials(pApPartials, nPartials, sh);
	if (pApStencilPartials) pAp += reduce_partials(pApStencilPartials, nPartials, sh);
	const double alphaD = st->rz / pAp;
	const float alpha = (float)alphaD;
	if (blockIdx.x == 0 && threadIdx.x == 0) stw->alpha = alphaD;
	double rr = 0.0;
	while (i0 < nv)
	{
#pragma unroll
		for (int u = 0; u < kVecUnroll; ++u)
		{
			const int i = i0 + u * stride;
			if (i >= nv) break;
			rv[u].x = fmaf(-alpha, av[u].x, rv[u].x); rv[u].y = fmaf(-alpha, av[u].y, rv[u].y); rv[u].z = fmaf(-alpha, av[u].z, rv[u].z);
			r[i] = rv[u];
			rr += (double)rv[u].x * rv[u].x + (double)rv[u].y * rv[u].y + (double)rv[u].z * rv[u].z;
		}
		i0 += kVecUnroll * stride;
#pragma unroll
		for (int u = 0; u < kVecUnroll; ++u)
		{
			const int i = i0 + u * stride;
			if (i < nv) { rv[u] = r[i]; av[u] = Ap[i]; }
		}
	}
	const double t = block_sum(rr, sh);
	if (threadIdx.x == 0) rrPartials[blockIdx.x] = t;
}

// partial sums of a.b (used for r.z every iteration and for the initial r.r / r.z)
__global__ void __launch_bounds__(kPcgThreads) dot_kernel(const float4* __restrict__ a, const float4* __restrict__ b, int nv,
	double* __restrict__ partials, const volatile PcgState* st)
{
	__shared__ double sh[kPcgWarps];
	if (st->done) return;
	double v = 0.0;
	const int stride = gridDim.x * blockDim.x;
	for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nv; i0 += kVecUnroll * stride)
	{
		float4 av[kVecUnroll], bv[kVecUnroll];
#pragma unroll
		for (int u = 0; u < kVecUnroll; ++u)
		{
			const int i = i0 + u * stride;
			if (i < nv) { av[u] = a[i]; bv[u] = b[i]; }
		}
#pragma unroll
		for (int u = 0; u < kVecUnroll; ++u)
		{
			if (i0 + u * stride >= nv) break;
			v += (double)av[u].x * bv[u].x + (double)av[u].y * bv[u].y + (double)av[u].z * bv[u].z;
		}
	}
	const double t = block_sum(v, sh);
	if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// mode 0 (setup): rr0 = rr = sum(rrPartials), rz = sum(rzPartials), p = z, it = 0
// mode 1 (iteration): x += alpha p, beta = rz'/rz, p = z + beta p; the last CTA records rr, the stopping test and the
// iteration count.  The flag written here is read only by LATER launches, so all CTAs of one launch see the same value.
__global__ void __launch_bounds__(kPcgThreads) update_p_kernel(float4* __restrict__ x, float4* __restrict__ p, const float4* __restrict__ z,
	int nv, const double* __restrict__ rzPartials, const double* __restrict__ rrPartials, int nPartials, double tol2, int maxIter,
	int mode, PcgState* stw)
{
	__shared__ double sh[2 * kPcgWarps];
	volatile PcgState* st = stw;
	if (st->done) return;
	const int stride = gridDim.x * blockDim.x;
	int i0 = blockIdx.x * blockDim.x + threadIdx.x;
	float4 zv[kVecUnroll], pv[kVecUnroll], xv[kVecUnroll];
	const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
	for (int u = 0; u < kVecUnroll; ++u)   // first trip's loads before the reduction
	{
		const int i = i0 + u * stride;
		if (i < nv) { zv[u] = z[i]; pv[u] = mode == 0 ? zero : p[i]; xv[u] = mode == 0 ? zero : x[i]; }
	}
	double rzNew, rr;
	reduce_partials2(rzPartials, rrPartials, nPartials, sh, rzNew, rr);
	const double rzOld = st->rz;
	const float beta = mode == 0 ? 0.f : (float)(rzNew / rzOld);
	const float alpha = mode == 0 ? 0.f : (float)st->alpha;
	while (i0 < nv)
	{
#pragma unroll
		for (int u = 0; u < kVecUnroll; ++u)
		{
			const int i = i0 + u * stride;
			if (i >= nv) break;
			if (mode != 0)
			{
				xv[u].x = fmaf(alpha, pv[u].x, xv[u].x); xv[u].y = fmaf(alpha, pv[u].y, xv[u].y); xv[u].z = fmaf(alpha, pv[u].z, xv[u].z);
				x[i] = xv[u];
			}
			pv[u].x = fmaf(beta, pv[u].x, zv[u].x); pv[u].y = fmaf(beta, pv[u].y, zv[u].y); pv[u].z = fmaf(beta, pv[u].z, zv[u].z);
			pv[u].w = 0.f;
			p[i] = pv[u];
		}
		i0 += kVecUnroll * stride;
#pragma unroll
		for (int u = 0; u < kVecUnroll; ++u)
		{
			const int i = i0 + u * stride;
			if (i < nv) { zv[u] = z[i]; pv[u] = mode == 0 ? zero : p[i]; xv[u] = mode == 0 ? zero : x[i]; }
		}
	}
	// grid-wide agreement: the state is rewritten by the LAST CTA to finish reading it
	__shared__ bool last;
	__threadfence();
	if (threadIdx.x == 0) last = atomicAdd(&stw->pad, 1) == (int)gridDim.x - 1;
	__syncthreads();
	if (last && threadIdx.x == 0)
	{
		st->pad = 0;
		st->rz = rzNew;
		st->rr = rr;
		if (mode == 0) { st->rr0 = rr; st->it = 0; st->iters = 0; if (rr == 0.0) st->done = 1; }
		else
		{
			st->it += 1;
			st->iters = st->it;
			if (rr < tol2 * st->rr0) st->done = 1;
			else if (st->it >= maxIter) st->done = 2;
		}
		__threadfence();
	}
}

__global__ void copy_b_kernel(const float4* __restrict__ b, float4* __restrict__ r, float4* __restrict__ x, int nv)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nv) return;
	const float4 v = b[i];
	r[i] = make_float4(v.x, v.y, v.z, 0.f);
	x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

#ifndef MAS_CPU_EMULATION   // host side: launches (tests/emu/pcg_emu.cpp, test infrastructure, has its own launcher)
int pcg_solve(Context* c, const float* diag, const float* off, const int* ranges, const int* idx, const float4* b, float4* x,
	float relTol, int maxIter, int usePrecond, int* itersOut, float* relResOut)
{
	cudaStream_t st = c->stream;
	const int nv = c->nv;
	if (int rc = reserve(c, c->pcgR, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgZ, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgP, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgAp, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgPartials, (size_t)4 * kMaxPartials)) return rc;
	if (int rc = reserve(c, c->pcgState, (size_t)sizeof(PcgState))) return rc;
	double* pA = c->pcgPartials.p;
	double* pRR = pA + kMaxPartials;
	double* pRZ = pRR + kMaxPartials;
	double* pS = pRZ + kMaxPartials;      // p.Ap, collision part
	// the collision stencils of the last prepare belong to A (the preconditioner was built for A + their Hessians)
	const int nStencil = c->prepared ? c->nStencil : 0;
	int gridStencil = cdiv(nStencil, kPcgThreads);
	if (gridStencil > kMaxPartials) gridStencil = kMaxPartials;
	PcgState* state = reinterpret_cast<PcgState*>(c->pcgState.p);
	float4 *r = c->pcgR.p, *z = usePrecond ? c->pcgZ.p : c->pcgR.p, *p = c->pcgP.p, *Ap = c->pcgAp.p;

	// resident grids: occupancy x SM count, never more CTAs than there is work for
	auto resident = [&](const void* kernel, long long workCtas) -> int {
		int perSm = 1;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kPcgThreads, 0) != cudaSuccess || perSm < 1) perSm = 1;
		long long g = (long long)perSm * c->smCount;
		if (g > workCtas) g = workCtas;
		if (g > kMaxPartials) g = kMaxPartials;
		return g < 1 ? 1 : (int)g;
	};
	const long long vecCtas = cdiv(nv, kPcgThreads);
	const int gridAxpy = resident((const void*)axpy_rr_kernel, vecCtas);
	const int gridDot = resident((const void*)dot_kernel, vecCtas);
	const int gridUpdate = resident((const void*)update_p_kernel, vecCtas);
	int gridSpmv = cdiv(cdiv(nv, 32), kSpmvWarps);          // one CTA per SM: its buffers take 180 KB of shared memory
	if (gridSpmv > c->smCount) gridSpmv = c->smCount;
	MAS_CUDA(c, cudaFuncSetAttribute((const void*)spmv_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes));
	// every partial array is summed over `nPart` entries: zeroed once, and each array is always filled by the same grid
	// (p.Ap: gridSpmv, r.r: gridAxpy, r.z: gridDot, collision part of p.Ap: gridStencil), so no stale entry survives a pass
	int nPart = gridAxpy;
	for (int g : { gridDot, gridUpdate, gridSpmv, gridStencil }) nPart = g > nPart ? g : nPart;
	const double tol2 = (double)relTol * (double)relTol;

	// sliced-ELL copy of A (see ell_fill_kernel): widths, a scan (its last entry is the total) and one pass over the blocks
	const int nSlices = cdiv(nv, 32);
	if (int rc = reserve(c, c->pcgSliceSlots, (size_t)nSlices + 1)) return rc;
	if (int rc = reserve(c, c->pcgSliceStart, (size_t)nSlices + 1)) return rc;
	if (int rc = reserve(c, c->scanTotal, 1)) return rc;
	ell_width_kernel<<<cdiv(((long long)nSlices + 1) * 32, 256), 256, 0, st>>>(ranges, nv, nSlices, c->pcgSliceSlots.p);
	if (int rc = launch_exclusive_scan(c, c->pcgSliceSlots.p, nSlices + 1, c->pcgSliceStart.p, c->scanTotal.p)) return rc;
	int totalSlots = 0;
	MAS_CUDA(c, cudaMemcpyAsync(&totalSlots, c->scanTotal.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	MAS_CUDA(c, cudaStreamSynchronize(st));
	if (int rc = reserve(c, c->pcgEllVal, (size_t)(totalSlots > 0 ? totalSlots : 1) * kSlotWords)) return rc;
	ell_fill_kernel<<<cdiv((long long)nSlices * 32, 256), 256, 0, st>>>(diag, off, ranges, idx, nv, c->pcgSliceStart.p, c->pcgEllVal.p);

	MAS_CUDA(c, cudaMemsetAsync(state, 0, sizeof(PcgState), st));
	MAS_CUDA(c, cudaMemsetAsync(pA, 0, sizeof(double) * 4 * kMaxPartials, st));
	copy_b_kernel<<<cdiv(nv, 256), 256, 0, st>>>(b, r, x, nv);
	auto precondition = [&]() -> int {
		if (!usePrecond) return MAS_OK;
		if (int rc = apply_begin(c, r)) return rc;
		return apply_end(c, r, z);
	};
	if (int rc = precondition()) return rc;
	dot_kernel<<<gridAxpy, kPcgThreads, 0, st>>>(r, r, nv, pRR, state);   // pRR always holds gridAxpy partials (axpy_rr refills it)
	dot_kernel<<<gridDot, kPcgThreads, 0, st>>>(r, z, nv, pRZ, state);
	update_p_kernel<<<gridUpdate, kPcgThreads, 0, st>>>(x, p, z, nv, pRZ, pRR, nPart, tol2, maxIter, 0, state);
	MAS_CUDA(c, cudaGetLastError());

	// one iteration, captured once
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	cudaStream_t cap;
	MAS_CUDA(c, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
	cudaStream_t saved = c->stream;
	c->stream = cap;
	int rc = MAS_OK;
	const int savedLaunches = c->applyLaunches;
	if (!check(c, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) rc = MAS_ERR_CUDA;
	if (rc == MAS_OK)
	{
		spmv_dot_kernel<<<gridSpmv, kSpmvThreads, kSpmvSmemBytes, cap>>>(c->pcgSliceStart.p, c->pcgEllVal.p, p, Ap, nv, pA, state,
			&state->stageErr);
		if (nStencil > 0)
			stencil_spmv_kernel<<<gridStencil, kPcgThreads, 0, cap>>>(c->stencils.p, nStencil, p, Ap, pS, state);
		axpy_rr_kernel<<<gridAxpy, kPcgThreads, 0, cap>>>(r, Ap, nv, pA, nStencil > 0 ? pS : nullptr, nPart, pRR, state);
		c->applyLaunches = 0;
		if (usePrecond) rc = apply_forked(c, r, z, cap);   // coarse chain concurrent with the head of the fine solve
		dot_kernel<<<gridDot, kPcgThreads, 0, cap>>>(r, z, nv, pRZ, state);
		update_p_kernel<<<gridUpdate, kPcgThreads, 0, cap>>>(x, p, z, nv, pRZ, pRR, nPart, tol2, maxIter, 1, state);
	}
	c->pcgLaunchesPerIter = 4 + (nStencil > 0 ? 1 : 0) + (usePrecond ? c->applyLaunches : 0);
	c->applyLaunches = savedLaunches;
	cudaError_t e = cudaStreamEndCapture(cap, &graph);
	c->stream = saved;
	if (rc == MAS_OK && !check(c, e, "cudaStreamEndCapture")) rc = MAS_ERR_CUDA;
	// same launch priorities as the apply's own graph: the coarse chain ahead of the streaming kernels it runs beside
	if (rc == MAS_OK && usePrecond) rc = prioritize_apply_graph(c, graph);
	if (rc == MAS_OK && !check(c, cudaGraphInstantiate(&exec, graph, 0), "cudaGraphInstantiate")) rc = MAS_ERR_CUDA;
	if (graph) cudaGraphDestroy(graph);
	cudaStreamDestroy(cap);
	if (rc != MAS_OK) return rc;

	PcgState host = {};
	int launched = 0;
	const int batch = 16;
	while (launched < maxIter)
	{
		const int n = (maxIter - launched) < batch ? (maxIter - launched) : batch;
		for (int k = 0; k < n; ++k)
			if (!check(c, cudaGraphLaunch(exec, st), "cudaGraphLaunch")) { cudaGraphExecDestroy(exec); return MAS_ERR_CUDA; }
		launched += n;
		if (!check(c, cudaMemcpyAsync(&host, state, sizeof(PcgState), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync") ||
			!check(c, cudaStreamSynchronize(st), "cudaStreamSynchronize")) { cudaGraphExecDestroy(exec); return MAS_ERR_CUDA; }
		if (host.stageErr)
		{
			cudaGraphExecDestroy(exec);
			c->err = "mas_pcg_solve: a bulk copy of the SpMV did not complete";
			return MAS_ERR_CUDA;
		}
		if (host.done) break;
	}
	if (launched == 0)
	{
		MAS_CUDA(c, cudaMemcpyAsync(&host, state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
		MAS_CUDA(c, cudaStreamSynchronize(st));
	}
	cudaGraphExecDestroy(exec);
	if (itersOut) *itersOut = host.iters;
	if (relResOut) *relResOut = host.rr0 > 0.0 ? (float)sqrt(host.rr / host.rr0) : 0.f;
	c->pcgConverged = host.done == 1;
	return MAS_OK;
}

#endif  // MAS_CPU_EMULATION

}  // namespace mas
