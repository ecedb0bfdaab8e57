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
// batch of iterations, not once per iteration.
//   spmv_dot      Ap = A p (warp per 32 rows, lane per 3x3 block, blocks staged through shared memory so every global
//                 load is coalesced; segmented shuffle scan adds the blocks of a row), partial sums of p.Ap
//   axpy_rr       alpha = rz / p.Ap;  x += alpha p;  r -= alpha Ap;  partial sums of r.r
//   (apply)       z = M^-1 r
//   dot_rz        partial sums of r.z; one thread evaluates the stopping test for this iteration
//   update_p      beta = rz' / rz;  p = z + beta p
#include "mas_internal.h"

namespace mas {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kPcgThreads = 256;
constexpr int kPcgWarps = kPcgThreads / 32;
constexpr int kMaxPartials = 1024;  // CTAs per reduction pass (grid-stride beyond that)

// Scalars: [0] rz  [1] pAp (unused, kept in partials)  [2] rr  [3] rr0  [4] rzNew
struct PcgState
{
	double rz, rr, rr0, rzNew;
	int done, iters, it;
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

// every CTA reduces the (<= kMaxPartials) partial sums of the previous pass in the same fixed order
__device__ __forceinline__ double reduce_partials(const double* __restrict__ partials, int n, double* sh)
{
	double v = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
	return block_sum(v, sh);
}

// Ap = A p and partial p.Ap.  One warp per 32 consecutive rows.
__global__ void __launch_bounds__(kPcgThreads) spmv_dot_kernel(const float* __restrict__ diag, const float* __restrict__ off,
	const int* __restrict__ ranges, const int* __restrict__ idx, const float4* __restrict__ p, float4* __restrict__ Ap, int nv,
	double* __restrict__ partials, const volatile PcgState* st)
{
	__shared__ float stage[kPcgWarps][288];
	__shared__ float acc[kPcgWarps][32][3];
	__shared__ double sh[kPcgWarps];
	if (st->done) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nWarpRows = (nv + 31) >> 5;
	double dot = 0.0;
	for (int wr = blockIdx.x * kPcgWarps + warp; wr < nWarpRows; wr += gridDim.x * kPcgWarps)
	{
		const int r0 = wr * 32, row = r0 + lane;
		const int rs = ranges[row < nv ? row : nv];   // start of my row (lanes past the end see the total)
		const int s = __shfl_sync(kFull, rs, 0);
		const int e = ranges[(r0 + 32) < nv ? (r0 + 32) : nv];
		acc[warp][lane][0] = acc[warp][lane][1] = acc[warp][lane][2] = 0.f;
		__syncwarp();
		for (int c = s; c < e; c += 32)
		{
			const int nb = min(32, e - c);
			// stage 32 blocks (288 floats) with coalesced loads
			const float* src = off + 9 * (size_t)c;
			for (int q = lane; q < 9 * nb; q += 32) stage[warp][q] = src[q];
			__syncwarp();
			const int blk = c + lane;
			const bool valid = lane < nb;
			float cx = 0.f, cy = 0.f, cz = 0.f;
			int key = -1;
			if (valid)
			{
				const float* m = &stage[warp][9 * lane];   // column-major: m[3j+i] = (i,j)
				const float4 xv = p[idx[blk]];
				cx = fmaf(m[0], xv.x, fmaf(m[3], xv.y, m[6] * xv.z));
				cy = fmaf(m[1], xv.x, fmaf(m[4], xv.y, m[7] * xv.z));
				cz = fmaf(m[2], xv.x, fmaf(m[5], xv.y, m[8] * xv.z));
			}
			// row of my block: the last lane l with rs_l <= blk (binary search over the warp's row starts)
			{
				int lo = 0, hi = 31;
				for (int it = 0; it < 5; ++it)
				{
					const int mid = (lo + hi + 1) >> 1;
					const int v = __shfl_sync(kFull, rs, mid);
					if (v <= blk) lo = mid; else hi = mid - 1;
				}
				if (valid) key = lo;
			}
			// segmented inclusive scan over lanes with equal key (contiguous), fixed order
			for (int o = 1; o < 32; o <<= 1)
			{
				const float ux = __shfl_up_sync(kFull, cx, o), uy = __shfl_up_sync(kFull, cy, o), uz = __shfl_up_sync(kFull, cz, o);
				const int uk = __shfl_up_sync(kFull, key, o);
				if (lane >= o && uk == key) { cx += ux; cy += uy; cz += uz; }
			}
			const int nk = __shfl_down_sync(kFull, key, 1);
			if (valid && (lane == 31 || nk != key))
			{
				acc[warp][key][0] += cx; acc[warp][key][1] += cy; acc[warp][key][2] += cz;
			}
			__syncwarp();
		}
		if (row < nv)
		{
			const float* d = diag + 9 * (size_t)row;
			const float4 xv = p[row];
			const float yx = fmaf(d[0], xv.x, fmaf(d[3], xv.y, fmaf(d[6], xv.z, acc[warp][lane][0])));
			const float yy = fmaf(d[1], xv.x, fmaf(d[4], xv.y, fmaf(d[7], xv.z, acc[warp][lane][1])));
			const float yz = fmaf(d[2], xv.x, fmaf(d[5], xv.y, fmaf(d[8], xv.z, acc[warp][lane][2])));
			Ap[row] = make_float4(yx, yy, yz, 0.f);
			dot += (double)xv.x * yx + (double)xv.y * yy + (double)xv.z * yz;
		}
		__syncwarp();
	}
	const double t = block_sum(dot, sh);
	if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kPcgThreads) axpy_rr_kernel(float4* __restrict__ x, float4* __restrict__ r,
	const float4* __restrict__ p, const float4* __restrict__ Ap, int nv, const double* __restrict__ pApPartials, int nPartials,
	double* __restrict__ rrPartials, const volatile PcgState* st)
{
	__shared__ double sh[kPcgWarps];
	if (st->done) return;
	const double pAp = reduce_partials(pApPartials, nPartials, sh);
	const float alpha = (float)(st->rz / pAp);
	double rr = 0.0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
	{
		float4 xv = x[i], rv = r[i];
		const float4 pv = p[i], av = Ap[i];
		xv.x = fmaf(alpha, pv.x, xv.x); xv.y = fmaf(alpha, pv.y, xv.y); xv.z = fmaf(alpha, pv.z, xv.z);
		rv.x = fmaf(-alpha, av.x, rv.x); rv.y = fmaf(-alpha, av.y, rv.y); rv.z = fmaf(-alpha, av.z, rv.z);
		x[i] = xv; r[i] = rv;
		rr += (double)rv.x * rv.x + (double)rv.y * rv.y + (double)rv.z * rv.z;
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
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
	{
		const float4 av = a[i], bv = b[i];
		v += (double)av.x * bv.x + (double)av.y * bv.y + (double)av.z * bv.z;
	}
	const double t = block_sum(v, sh);
	if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// mode 0 (setup): rr0 = rr = sum(rrPartials), rz = sum(rzPartials), p = z, it = 0
// mode 1 (iteration): beta = rz'/rz, p = z + beta p; block 0 records rr, the stopping test and the iteration count.
// The flag written here is read only by LATER launches, so all CTAs of one launch see the same value.
__global__ void __launch_bounds__(kPcgThreads) update_p_kernel(float4* __restrict__ p, const float4* __restrict__ z, int nv,
	const double* __restrict__ rzPartials, const double* __restrict__ rrPartials, int nPartials, double tol2, int mode,
	PcgState* stw)
{
	__shared__ double sh[kPcgWarps];
	volatile PcgState* st = stw;
	if (st->done) return;
	const double rzNew = reduce_partials(rzPartials, nPartials, sh);
	const double rr = reduce_partials(rrPartials, nPartials, sh);
	const double rzOld = st->rz;
	const float beta = mode == 0 ? 0.f : (float)(rzNew / rzOld);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
	{
		const float4 zv = z[i];
		float4 pv = mode == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : p[i];
		pv.x = fmaf(beta, pv.x, zv.x); pv.y = fmaf(beta, pv.y, zv.y); pv.z = fmaf(beta, pv.z, zv.z); pv.w = 0.f;
		p[i] = pv;
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

int pcg_solve(Context* c, const float* diag, const float* off, const int* ranges, const int* idx, const float4* b, float4* x,
	float relTol, int maxIter, int usePrecond, int* itersOut, float* relResOut)
{
	cudaStream_t st = c->stream;
	const int nv = c->nv;
	if (int rc = reserve(c, c->pcgR, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgZ, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgP, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgAp, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->pcgPartials, (size_t)3 * kMaxPartials)) return rc;
	if (int rc = reserve(c, c->pcgState, (size_t)sizeof(PcgState))) return rc;
	double* pA = c->pcgPartials.p;
	double* pRR = pA + kMaxPartials;
	double* pRZ = pRR + kMaxPartials;
	PcgState* state = reinterpret_cast<PcgState*>(c->pcgState.p);
	float4 *r = c->pcgR.p, *z = usePrecond ? c->pcgZ.p : c->pcgR.p, *p = c->pcgP.p, *Ap = c->pcgAp.p;

	int grid = cdiv(nv, kPcgThreads);
	if (grid > kMaxPartials) grid = kMaxPartials;
	int gridSpmv = cdiv(cdiv(nv, 32), kPcgWarps);
	if (gridSpmv > kMaxPartials) gridSpmv = kMaxPartials;
	// all three partial arrays are summed over `nPart` entries: zero them once, passes fill what they use
	const int nPart = grid > gridSpmv ? grid : gridSpmv;
	const double tol2 = (double)relTol * (double)relTol;

	MAS_CUDA(c, cudaMemsetAsync(state, 0, sizeof(PcgState), st));
	MAS_CUDA(c, cudaMemsetAsync(pA, 0, sizeof(double) * 3 * kMaxPartials, st));
	copy_b_kernel<<<cdiv(nv, 256), 256, 0, st>>>(b, r, x, nv);
	auto precondition = [&]() -> int {
		if (!usePrecond) return MAS_OK;
		if (int rc = apply_begin(c, r)) return rc;
		return apply_end(c, r, z);
	};
	if (int rc = precondition()) return rc;
	dot_kernel<<<grid, kPcgThreads, 0, st>>>(r, r, nv, pRR, state);
	dot_kernel<<<grid, kPcgThreads, 0, st>>>(r, z, nv, pRZ, state);
	update_p_kernel<<<grid, kPcgThreads, 0, st>>>(p, z, nv, pRZ, pRR, nPart, tol2, 0, state);
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
		spmv_dot_kernel<<<gridSpmv, kPcgThreads, 0, cap>>>(diag, off, ranges, idx, p, Ap, nv, pA, state);
		axpy_rr_kernel<<<grid, kPcgThreads, 0, cap>>>(x, r, p, Ap, nv, pA, nPart, pRR, state);
		c->applyLaunches = 0;
		if (usePrecond) rc = apply_forked(c, r, z, cap);   // coarse chain concurrent with the head of the fine solve
		dot_kernel<<<grid, kPcgThreads, 0, cap>>>(r, z, nv, pRZ, state);
		update_p_kernel<<<grid, kPcgThreads, 0, cap>>>(p, z, nv, pRZ, pRR, nPart, tol2, 1, state);
	}
	c->pcgLaunchesPerIter = 4 + (usePrecond ? c->applyLaunches : 0);
	c->applyLaunches = savedLaunches;
	cudaError_t e = cudaStreamEndCapture(cap, &graph);
	c->stream = saved;
	if (rc == MAS_OK && !check(c, e, "cudaStreamEndCapture")) rc = MAS_ERR_CUDA;
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
	c->pcgConverged = host.done;
	return MAS_OK;
}

}  // namespace mas
