// Device-resident preconditioned conjugate gradients: the immediate CALLER of Preconditioning().
// The reference ships no solver (its caller owns the PCG loop, SeSchwarzPreconditioner.h:55-63); BASELINE config 2
// ("PCG to 1e-5 residual with MAS, iteration count and wall time") needs one on both sides, and without it the
// host<->device copies of r and z would dwarf the apply.
//
//   A is the caller's Hessian in the same arrays PreparePreconditioner receives (original vertex order):
//   diag[nv] + offdiag[nnz] column-major 3x3 blocks, CSR adjacency (ranges, idx); vectors are 16-byte xyzw.
//   x0 = 0;  stop when ||r||_2 / ||b||_2 < relTol;  dot products in FP64 with a fixed reduction order.
//
// One iteration = 4 launches + the apply's own launches; kIterPerGraph iterations are captured once in a CUDA graph and
// replayed.  The stopping test and the iteration limit are evaluated on the device (a flag turns the remaining launches into
// no-ops), so the host synchronises once per batch of graph launches, not once per iteration.  Every kernel is launched with exactly as many CTAs as are resident at once
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

#include <vector>

namespace mas {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kPcgThreads = 256;
constexpr int kPcgWarps = kPcgThreads / 32;
constexpr int kMaxPartials = 2048;  // upper bound of CTAs per reduction pass (the grids are occupancy x SM count)
constexpr int kSpmvBatch = 4;       // block slots whose loads are in flight together in the SpMV
constexpr int kIterPerGraph = 1;    // PCG iterations per captured graph (4: 203.4 vs 204.8 us per iteration, but up to three idle
                                    // iterations - each still runs the apply - at the end of every solve: 20.1 vs 19.9 ms at 1M vertices)
constexpr int kMaxBatch = 16;       // graph launches between two looks at the state
constexpr int kVecUnroll = 4;       // elements per thread and trip of the vector kernels, loads in flight together

struct PcgState
{
	double rz, rr, rr0, alpha;   // alpha: step length of the running iteration (axpy_rr -> update_p)
	int done, iters, it;         // done: 1 = converged, 2 = maxIter reached
	int pad;
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

// the ELL image is read once per SpMV: no L1 allocation, first in line for eviction from L2 (the vectors of the iteration,
// 16 MB each at 1M vertices, are what should stay there between the kernels)
#ifdef MAS_CPU_EMULATION
__device__ __forceinline__ float ell_load(const float* p) { return *p; }
__device__ __forceinline__ int ell_load(const int* p) { return *p; }
#else
#ifndef MAS_L2_STREAM_HINT
#define MAS_L2_STREAM_HINT 1
#endif
__device__ __forceinline__ float ell_load(const float* p)
{
#if MAS_L2_STREAM_HINT
	float v;
	unsigned long long pol;
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
	return v;
#else
	return *p;
#endif
}
__device__ __forceinline__ int ell_load(const int* p) { return __float_as_int(ell_load(reinterpret_cast<const float*>(p))); }
#endif

// every CTA reduces the (<= kMaxPartials) partial sums of the previous pass in the same fixed order
__device__ __forceinline__ double reduce_partials(const double* __restrict__ partials, int n, double* sh)
{
	double v = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
	return block_sum(v, sh);
}

// ---- sliced-ELL copy of the caller's block CSR (once per solve) ---------------------------------------------------------
// slice g = rows 32g .. 32g+31, width w_g = 1 + its longest row; slot (g, k, lane) holds the diagonal block of row 32g+lane
// for k = 0 and the (k-1)-th off-diagonal block of that row after it:
//   ellIdx[sliceStart[g] + 32 k + lane]                (column vertex, -1 = padding)
//   ellVal[9 (sliceStart[g] + 32 k) + 32 e + lane]     (entry e of the column-major 3x3 block)
__global__ void ell_width_kernel(const int* __restrict__ ranges, int nv, int* __restrict__ sliceSlots)
{
	const int lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (g * 32 >= nv) return;
	const int row = g * 32 + lane;
	int deg = row < nv ? ranges[row + 1] - ranges[row] : 0;
	for (int off = 16; off > 0; off >>= 1) deg = max(deg, __shfl_xor_sync(kFull, deg, off));
	if (lane == 0) sliceSlots[g] = 32 * (deg + 1);
}

__global__ void ell_fill_kernel(const float* __restrict__ diag, const float* __restrict__ off, const int* __restrict__ ranges,
	const int* __restrict__ idx, int nv, const int* __restrict__ sliceStart, const int* __restrict__ sliceSlots, int* __restrict__ ellIdx,
	float* __restrict__ ellVal)
{
	const int lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (g * 32 >= nv) return;
	const int row = g * 32 + lane;
	const int rs = row < nv ? ranges[row] : 0, re = row < nv ? ranges[row + 1] : 0;
	const int base = sliceStart[g], width = sliceSlots[g] >> 5;
	for (int k = 0; k < width; ++k)
	{
		const bool has = k == 0 ? row < nv : rs + k - 1 < re;
		ellIdx[base + 32 * k + lane] = !has ? -1 : k == 0 ? row : idx[rs + k - 1];
		const float* m = k == 0 ? diag + 9 * (size_t)row : off + 9 * (size_t)(rs + k - 1);
		float* dst = ellVal + 9 * (size_t)(base + 32 * k) + lane;
#pragma unroll
		for (int e = 0; e < 9; ++e) dst[32 * e] = has ? m[e] : 0.0f;
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

// Ap = A p and partial p.Ap.  One warp per 32-row slice at a time, lane = row; the warps of the (resident) grid stride over
// the slices.
__global__ void __launch_bounds__(kPcgThreads) spmv_dot_kernel(const int* __restrict__ sliceStart,
	const int* __restrict__ sliceSlots, const int* __restrict__ ellIdx, const float* __restrict__ ellVal, const float4* __restrict__ p,
	float4* __restrict__ Ap, int nv, double* __restrict__ partials, const volatile PcgState* st)
{
	__shared__ double sh[kPcgWarps];
	if (st->done) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nSlices = (nv + 31) >> 5;
	double dot = 0.0;
	for (int g = blockIdx.x * kPcgWarps + warp; g < nSlices; g += gridDim.x * kPcgWarps)
	{
		const int row = g * 32 + lane;
		const int base = sliceStart[g], width = sliceSlots[g] >> 5;
		float yx = 0.f, yy = 0.f, yz = 0.f;
		// Branch-free and in batches of kSpmvBatch slots: first every load of the batch (column indices and blocks, then the
		// gathers of p), then the FMAs.  Padding slots carry zero blocks and read p[0].  With a branch on the column index
		// every slot cost two dependent memory round trips and the kernel ran at 3.6 TB/s.
		for (int k0 = 0; k0 < width; k0 += kSpmvBatch)
		{
			int c[kSpmvBatch];
			float v[kSpmvBatch][9];
			float4 q[kSpmvBatch];
#pragma unroll
			for (int u = 0; u < kSpmvBatch; ++u)
			{
				const bool in = k0 + u < width;
				const int slot = base + 32 * (in ? k0 + u : k0);
				c[u] = in ? ell_load(ellIdx + slot + lane) : -1;
				const float* m = ellVal + 9 * (size_t)slot + lane;
#pragma unroll
				for (int e = 0; e < 9; ++e) v[u][e] = in ? ell_load(m + 32 * e) : 0.0f;
			}
#pragma unroll
			for (int u = 0; u < kSpmvBatch; ++u) q[u] = p[c[u] < 0 ? 0 : c[u]];
#pragma unroll
			for (int u = 0; u < kSpmvBatch; ++u)
			{
				yx = fmaf(v[u][0], q[u].x, fmaf(v[u][3], q[u].y, fmaf(v[u][6], q[u].z, yx)));
				yy = fmaf(v[u][1], q[u].x, fmaf(v[u][4], q[u].y, fmaf(v[u][7], q[u].z, yy)));
				yz = fmaf(v[u][2], q[u].x, fmaf(v[u][5], q[u].y, fmaf(v[u][8], q[u].z, yz)));
			}
		}
		if (row < nv)
		{
			const float4 xv = p[row];   // the diagonal slot gathered it a moment ago
			Ap[row] = make_float4(yx, yy, yz, 0.f);
			dot += (double)xv.x * yx + (double)xv.y * yy + (double)xv.z * yz;
		}
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
	// r, z, p, Ap in ONE allocation: a single L2 access-policy window covers them (see below)
	const size_t nvPad = ((size_t)nv + 63) / 64 * 64;
	if (int rc = reserve(c, c->pcgR, 4 * nvPad)) return rc;
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
	float4 *r = c->pcgR.p, *z = usePrecond ? r + nvPad : r, *p = r + 2 * nvPad, *Ap = r + 3 * nvPad;

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
	const int gridSpmv = resident((const void*)spmv_dot_kernel, cdiv(cdiv(nv, 32), kPcgWarps));
	// every partial array is summed over `nPart` entries: zeroed once, and each array is always filled by the same grid
	// (p.Ap: gridSpmv, r.r: gridAxpy, r.z: gridDot, collision part of p.Ap: gridStencil), so no stale entry survives a pass
	int nPart = gridAxpy;
	for (int g : { gridDot, gridUpdate, gridSpmv, gridStencil }) nPart = g > nPart ? g : nPart;
	const double tol2 = (double)relTol * (double)relTol;

	// sliced-ELL copy of A (see ell_fill_kernel): two small passes, a scan and one pass over the blocks
	const int nSlices = cdiv(nv, 32);
	if (int rc = reserve(c, c->pcgSliceSlots, (size_t)nSlices)) return rc;
	if (int rc = reserve(c, c->pcgSliceStart, (size_t)nSlices)) return rc;
	if (int rc = reserve(c, c->scanTotal, 1)) return rc;
	ell_width_kernel<<<cdiv((long long)nSlices * 32, 256), 256, 0, st>>>(ranges, nv, c->pcgSliceSlots.p);
	if (int rc = launch_exclusive_scan(c, c->pcgSliceSlots.p, nSlices, c->pcgSliceStart.p, c->scanTotal.p)) return rc;
	int totalSlots = 0;
	MAS_CUDA(c, cudaMemcpyAsync(&totalSlots, c->scanTotal.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	MAS_CUDA(c, cudaStreamSynchronize(st));
	if (int rc = reserve(c, c->pcgEllIdx, (size_t)(totalSlots > 0 ? totalSlots : 1))) return rc;
	if (int rc = reserve(c, c->pcgEllVal, (size_t)(totalSlots > 0 ? totalSlots : 1) * 9)) return rc;
	ell_fill_kernel<<<cdiv((long long)nSlices * 32, 256), 256, 0, st>>>(diag, off, ranges, idx, nv, c->pcgSliceStart.p, c->pcgSliceSlots.p,
		c->pcgEllIdx.p, c->pcgEllVal.p);

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

	// kIterPerGraph iterations, captured once
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	cudaStream_t cap;
	MAS_CUDA(c, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
	cudaStream_t saved = c->stream;
	c->stream = cap;
	int rc = MAS_OK;
	const int savedLaunches = c->applyLaunches;
	if (!check(c, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) rc = MAS_ERR_CUDA;
	for (int k = 0; k < kIterPerGraph && rc == MAS_OK; ++k)
	{
		spmv_dot_kernel<<<gridSpmv, kPcgThreads, 0, cap>>>(c->pcgSliceStart.p, c->pcgSliceSlots.p, c->pcgEllIdx.p, c->pcgEllVal.p, p, Ap,
			nv, pA, state);
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
	// The four vectors of the iteration (64 MB at 1M vertices) are written by one kernel and read by the next while a
	// gigabyte of matrix and packed inverses streams through the 126 MB L2 in between.  A persisting access-policy window
	// over them on every kernel node (hits persist, everything else is treated as streaming) keeps them on chip; the
	// carve-out lasts for this solve only.  Only when all four fit the carve-out the device allows: a window larger than
	// that (hit ratio < 1) thrashes - 1,051 against 784 us per iteration on the 4.2M-vertex cloth, whose vectors take 256 MB -
	// while at 1M vertices the iteration drops from 212 to 205 us (profiles/r02_pcg_l2_persistence.txt).  Best effort: a
	// device that refuses any of it runs as before.
	bool persisting = false;
	size_t l2LimitBefore = 0;   // the application's own carve-out, restored when the solve returns
	if (rc == MAS_OK)
	{
		int maxPersist = 0, maxWindow = 0;
		cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
		cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
		const size_t arena = 4 * nvPad * sizeof(float4);
		const size_t window = arena, carve = arena <= (size_t)maxPersist && arena <= (size_t)maxWindow ? arena : 0;
		if (cudaDeviceGetLimit(&l2LimitBefore, cudaLimitPersistingL2CacheSize) != cudaSuccess) l2LimitBefore = 0;
		if (c->optPcgPersistL2 && carve > 0 &&
			cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve > l2LimitBefore ? carve : l2LimitBefore) == cudaSuccess)
		{
			persisting = true;
			cudaKernelNodeAttrValue v = {};
			v.accessPolicyWindow.base_ptr = r;
			v.accessPolicyWindow.num_bytes = window;
			v.accessPolicyWindow.hitRatio = (float)((double)carve / (double)window);
			v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			size_t n = 0;
			cudaGraphGetNodes(graph, nullptr, &n);
			std::vector<cudaGraphNode_t> nodes(n);
			if (n) cudaGraphGetNodes(graph, nodes.data(), &n);
			for (size_t i = 0; i < n; ++i)
			{
				cudaGraphNodeType type;
				if (cudaGraphNodeGetType(nodes[i], &type) == cudaSuccess && type == cudaGraphNodeTypeKernel)
					cudaGraphKernelNodeSetAttribute(nodes[i], cudaKernelNodeAttributeAccessPolicyWindow, &v);
			}
		}
		cudaGetLastError();
	}
	if (rc == MAS_OK && !check(c, cudaGraphInstantiate(&exec, graph, 0), "cudaGraphInstantiate")) rc = MAS_ERR_CUDA;
	if (graph) cudaGraphDestroy(graph);
	cudaStreamDestroy(cap);

	auto release_l2 = [&]() {
		if (!persisting) return;
		cudaCtxResetPersistingL2Cache();
		cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2LimitBefore);
		cudaGetLastError();
	};
	if (rc != MAS_OK) { release_l2(); return rc; }
	// The host looks at the state after every batch of launches.  Launches past the stopping test are no-ops in the PCG kernels
	// but still run the apply (104 us each at 1M vertices), so the batch shrinks as the end comes into sight: the iterations
	// left are extrapolated from the average contraction so far, and 60 % of them are launched.
	PcgState host = {};
	int launched = 0, batch = kMaxBatch < 8 ? kMaxBatch : 8;
	while (launched < maxIter)
	{
		for (int k = 0; k < batch && launched < maxIter; ++k, launched += kIterPerGraph)
			if (!check(c, cudaGraphLaunch(exec, st), "cudaGraphLaunch")) { cudaGraphExecDestroy(exec); release_l2(); return MAS_ERR_CUDA; }
		if (!check(c, cudaMemcpyAsync(&host, state, sizeof(PcgState), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync") ||
			!check(c, cudaStreamSynchronize(st), "cudaStreamSynchronize")) { cudaGraphExecDestroy(exec); release_l2(); return MAS_ERR_CUDA; }
		if (host.done) break;
		batch = kMaxBatch;
		if (host.it > 0 && host.rr > 0.0 && host.rr < host.rr0)
		{
			const double perIter = log(host.rr / host.rr0) / host.it;            // < 0
			const double left = (log(tol2) - log(host.rr / host.rr0)) / perIter;   // iterations until rr < tol2 rr0 at that rate
			const int want = (int)(0.6 * left / kIterPerGraph);
			batch = want < 1 ? 1 : want > kMaxBatch ? kMaxBatch : want;
		}
	}
	if (launched == 0)
	{
		MAS_CUDA(c, cudaMemcpyAsync(&host, state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
		MAS_CUDA(c, cudaStreamSynchronize(st));
	}
	cudaGraphExecDestroy(exec);
	release_l2();
	if (itersOut) *itersOut = host.iters;
	if (relResOut) *relResOut = host.rr0 > 0.0 ? (float)sqrt(host.rr / host.rr0) : 0.f;
	c->pcgConverged = host.done == 1;
	return MAS_OK;
}

#endif  // MAS_CPU_EMULATION

}  // namespace mas
