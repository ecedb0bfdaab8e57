// Vertex ordering: AABB, 63-bit Morton codes, sort, inverse permutation and the
// adjacency in sorted space.  Replaces the body of AllocatePrecoditioner
// (SeSchwarzPreconditioner.cpp:38-65): ComputeTotalAABB/ComputeAABB (193-211),
// FillSortingData (219-235) with SeMorton64::Encode (SeMorton.h:75-101),
// DoingSort (238-243), ComputeInverseMapper (245-255), MapHessianTable (258-285).
//
// Bit-exactness rules: IEEE round-to-nearest sub/div/mul (no contraction), the
// reference's comparison-based clamp (a NaN comes out as the upper bound), and
// truncating float->u64 conversion.  Equal codes are ordered by ascending
// original index (stable LSD radix sort over an iota payload); the reference's
// std::sort leaves that order unspecified.
#include "mas_internal.h"

#include <cfloat>
#ifndef MAS_CPU_EMULATION   // tests/emu/order_emu.cpp compiles the kernels of this file for the host (test infrastructure)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

namespace mas {

namespace {

constexpr int kReduceThreads = 256;

// SeAabbSimd.h:76-79 folds with _mm_min_ps/_mm_max_ps (a<b?a:b / a>b?a:b with a = running bound).
// For finite inputs min/max are order-independent, so a tree reduction gives the same bits.
__device__ __forceinline__ float4 vmin(float4 a, float4 b)
{
	return make_float4(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z, a.w < b.w ? a.w : b.w);
}
__device__ __forceinline__ float4 vmax(float4 a, float4 b)
{
	return make_float4(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z, a.w > b.w ? a.w : b.w);
}
__device__ __forceinline__ float4 shfl_down4(float4 v, int off)
{
	return make_float4(__shfl_down_sync(0xffffffffu, v.x, off), __shfl_down_sync(0xffffffffu, v.y, off),
		__shfl_down_sync(0xffffffffu, v.z, off), __shfl_down_sync(0xffffffffu, v.w, off));
}

// stage 1: per-CTA partial bounds; stage 2 (count == #partials, one CTA) folds them.
// `partials` holds lower bounds in [0, n) and upper bounds in [n, 2n).
__global__ void __launch_bounds__(kReduceThreads) aabb_partial_kernel(const float4* __restrict__ lowIn,
	const float4* __restrict__ highIn, int count, float4* __restrict__ partials, int nPartials)
{
	// Lower(FLT_MAX), Upper(-FLT_MAX) with w = 0 (SeAabbSimd.h:51 via SeVectorSimd.h:60)
	float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f);
	float4 hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
	{
		lo = vmin(lo, lowIn[i]);
		hi = vmax(hi, highIn[i]);
	}
	for (int off = 16; off > 0; off >>= 1)
	{
		lo = vmin(lo, shfl_down4(lo, off));
		hi = vmax(hi, shfl_down4(hi, off));
	}
	__shared__ float4 sLo[kReduceThreads / 32], sHi[kReduceThreads / 32];
	int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0) { sLo[warp] = lo; sHi[warp] = hi; }
	__syncthreads();
	if (warp == 0)
	{
		lo = lane < kReduceThreads / 32 ? sLo[lane] : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f);
		hi = lane < kReduceThreads / 32 ? sHi[lane] : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
		for (int off = 4; off > 0; off >>= 1)
		{
			lo = vmin(lo, shfl_down4(lo, off));
			hi = vmax(hi, shfl_down4(hi, off));
		}
		if (lane == 0)
		{
			partials[blockIdx.x] = lo;
			partials[nPartials + blockIdx.x] = hi;
		}
	}
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long b)
{
	b = (b | (b << 32)) & 0xFFFF00000000FFFFull;
	b = (b | (b << 16)) & 0x00FF0000FF0000FFull;
	b = (b | (b << 8)) & 0xF00F00F00F00F00Full;
	b = (b | (b << 4)) & 0x30C30C30C30C30C3ull;
	return (b | (b << 2)) & 0x9249249249249249ull;
}

// Math::Clamp(a, lo, hi) = Min(Max(lo, a), hi) with the ?: macros of SePreDefine.h:37-38
__device__ __forceinline__ unsigned long long axis_bits(float c)
{
	c = __fmul_rn(c, 2097152.0f);
	float t = (0.0f > c) ? 0.0f : c;
	t = (t < 2097151.0f) ? t : 2097151.0f;
	return spread3((unsigned long long)t);
}

__device__ __forceinline__ unsigned long long morton63(float x, float y, float z)
{
	return (axis_bits(x) << 2) + (axis_bits(y) << 1) + axis_bits(z);
}

__global__ void morton_kernel(const float4* __restrict__ pos, const float4* __restrict__ bounds, int nv,
	unsigned long long* __restrict__ code, int* __restrict__ iota)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nv) return;
	float4 lo = bounds[0], hi = bounds[1];
	float4 p = pos[v];
	// (m_positions[vid] - m_aabb.Lower) / m_aabb.Extent()   (cpp:225, subps/divps)
	float tx = __fdiv_rn(__fsub_rn(p.x, lo.x), __fsub_rn(hi.x, lo.x));
	float ty = __fdiv_rn(__fsub_rn(p.y, lo.y), __fsub_rn(hi.y, lo.y));
	float tz = __fdiv_rn(__fsub_rn(p.z, lo.z), __fsub_rn(hi.z, lo.z));
	code[v] = morton63(tx, ty, tz);
	iota[v] = v;
}

__global__ void morton_points_kernel(const float* __restrict__ xyz, int count, unsigned long long* __restrict__ out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) out[i] = morton63(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

__global__ void inverse_perm_kernel(const int* __restrict__ s2o, int nv, int* __restrict__ o2s)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v < nv) o2s[s2o[v]] = v;
}

__global__ void sorted_degree_kernel(const int* __restrict__ s2o, const int* __restrict__ inStarts, int nv, int* __restrict__ deg)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v > nv) return;
	if (v == nv) { deg[v] = 0; return; }
	int ov = s2o[v];
	deg[v] = inStarts[ov + 1] - inStarts[ov];
}

// m_mappedNeighbors[k][vid] = originalGetSorted[neighbors[k-1]] (cpp:278-283), stored as CSR rows in sorted space
__global__ void remap_adjacency_kernel(const int* __restrict__ s2o, const int* __restrict__ o2s,
	const int* __restrict__ inStarts, const int* __restrict__ inIdx, const int* __restrict__ adjStart, int nv,
	int* __restrict__ adjIdx)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nv) return;
	int ov = s2o[v];
	int src = inStarts[ov], n = inStarts[ov + 1] - src, dst = adjStart[v];
	for (int k = 0; k < n; ++k) adjIdx[dst + k] = o2s[inIdx[src + k]];
}

}  // namespace

#ifndef MAS_CPU_EMULATION   // host side: launches and the CUB sort / scan
int order_vertices(Context* c, const float4* positions, const int* inStarts, const int* inIdx)
{
	const int nv = c->nv;
	cudaStream_t s = c->stream;
	const int threads = 256;

	// ---- AABB
	int nPartials = c->smCount * 4;
	if (nPartials > cdiv(nv, kReduceThreads)) nPartials = cdiv(nv, kReduceThreads);
	if (nPartials < 1) nPartials = 1;
	TempBuf<float4> partials;
	if (int rc = reserve(c, partials, (size_t)2 * nPartials)) return rc;
	if (int rc = reserve(c, c->aabb, 8)) return rc;
	aabb_partial_kernel<<<nPartials, kReduceThreads, 0, s>>>(positions, positions, nv, partials.p, nPartials);
	aabb_partial_kernel<<<1, kReduceThreads, 0, s>>>(partials.p, partials.p + nPartials, nPartials, (float4*)c->aabb.p, 1);

	// ---- Morton codes + iota
	if (int rc = reserve(c, c->code, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->codeSorted, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->iota, (size_t)nv + 1)) return rc;
	if (int rc = reserve(c, c->s2o, (size_t)nv)) return rc;
	if (int rc = reserve(c, c->o2s, (size_t)nv)) return rc;
	morton_kernel<<<cdiv(nv, threads), threads, 0, s>>>(positions, (const float4*)c->aabb.p, nv, c->code.p, c->iota.p);

	// ---- sort (key = code, payload = original index); 63 significant bits
	size_t tempBytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, c->code.p, c->codeSorted.p, c->iota.p, c->s2o.p, nv, 0, 63, s);
	size_t scanBytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, c->iota.p, c->iota.p, nv + 1, s);
	if (scanBytes > tempBytes) tempBytes = scanBytes;
	if (int rc = reserve(c, c->cubTemp, tempBytes)) return rc;
	MAS_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, tempBytes, c->code.p, c->codeSorted.p, c->iota.p, c->s2o.p, nv, 0, 63, s));

	inverse_perm_kernel<<<cdiv(nv, threads), threads, 0, s>>>(c->s2o.p, nv, c->o2s.p);

	// ---- adjacency in sorted space
	if (int rc = reserve(c, c->adjStart, (size_t)nv + 1)) return rc;
	if (int rc = reserve(c, c->adjIdx, (size_t)(c->nnz > 0 ? c->nnz : 1))) return rc;
	sorted_degree_kernel<<<cdiv(nv + 1, threads), threads, 0, s>>>(c->s2o.p, inStarts, nv, c->iota.p);
	MAS_CUDA(c, cub::DeviceScan::ExclusiveSum(c->cubTemp.p, tempBytes, c->iota.p, c->adjStart.p, nv + 1, s));
	remap_adjacency_kernel<<<cdiv(nv, threads), threads, 0, s>>>(c->s2o.p, c->o2s.p, inStarts, inIdx, c->adjStart.p, nv, c->adjIdx.p);
	MAS_CUDA(c, cudaGetLastError());
	MAS_CUDA(c, cudaStreamSynchronize(s));
	return MAS_OK;
}

int morton_encode_points(Context* c, const float* xyz, int count, unsigned long long* out)
{
	TempBuf<float> in;
	TempBuf<unsigned long long> codes;
	if (int rc = reserve(c, in, (size_t)3 * count)) return rc;
	if (int rc = reserve(c, codes, (size_t)count)) return rc;
	MAS_CUDA(c, cudaMemcpyAsync(in.p, xyz, sizeof(float) * 3 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
	morton_points_kernel<<<cdiv(count, 256), 256, 0, c->stream>>>(in.p, count, codes.p);
	MAS_CUDA(c, cudaMemcpyAsync(out, codes.p, sizeof(unsigned long long) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
	MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	return MAS_OK;
}

#endif  // MAS_CPU_EMULATION

}  // namespace mas
