// Collision stencils and the multilevel clustering ("ReorderRealtime").
// Replaces PrepareCollisionStencils + MapCollisionStencilIndices
// (SeSchwarzPreconditioner.cpp:287-413) and ReorderRealtime (415-445):
// BuildConnectMaskL0 (447-511), BuildCollisionConnection (514-563),
// PreparePrefixSumL0 (565-628), BuildLevel1 (630-740), BuildConnectMaskLx
// (743-871), NextLevelCluster (873-961), PrefixSumLx (963-1072),
// ComputeNextLevel (1074-1084), TotalNodes (1086-1090), AggregationKernel
// (1092-1162).  All of it is integer work and must be bit-exact.
//
// One warp owns one 32-node bank; connectivity masks live in registers and the
// in-bank transitive closure runs on warp shuffles.  Cluster ids come from an
// exclusive scan of the per-bank counts of elected (lowest-lane) nodes.
//
// Differences from the reference, none of which change results on valid input:
//  * the "remaining neighbour" lists (cpp:74-75, 486-491, 788-793) are not kept:
//    an edge consumed at one level joins both ends in one cluster, so later it
//    could only set a node's own bit, which cpp:898 sets anyway;
//  * ids come from a correct scan (bug Q5 at cpp:989-994 is not reproduced);
//  * stencils are compacted in input order (the reference's order is
//    thread-timing dependent, cpp:407).
#include "mas_internal.h"
#ifndef MAS_CPU_EMULATION
#include <cub/device/device_scan.cuh>
#endif

#include <utility>

namespace mas {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned lanemask_lt(unsigned lane) { return (1u << lane) - 1u; }

// ------------------------------------------------------------------ stencils
__device__ __forceinline__ float ldf(const unsigned char* p, int off) { return *reinterpret_cast<const float*>(p + off); }
__device__ __forceinline__ int ldi(const unsigned char* p, int off) { return *reinterpret_cast<const int*>(p + off); }

// valid flag per candidate stencil (cpp:330, 359, 385).  An id beyond the mesh (the reference would read past its edge / face
// arrays) drops the stencil and raises *inputErr, which fails the prepare.
__global__ void stencil_flag_kernel(const unsigned char* __restrict__ ef, const unsigned char* __restrict__ ee,
	const unsigned char* __restrict__ vf, int efN, int eeN, int total, int fix, int nv, int ne, int nf, int* __restrict__ flag,
	int* __restrict__ inputErr)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total) return;
	// Q2: the reference reads every kind at the GLOBAL stencil index; MAS_OPT_STENCIL_FIX reads each array from its own zero
	const int local = i < efN ? i : (fix ? (i < efN + eeN ? i - efN : i - efN - eeN) : i);
	const unsigned char* rec = (i < efN ? ef : (i < efN + eeN ? ee : vf)) + 48 * (size_t)local;
	const int id0 = ldi(rec, 0), id1 = ldi(rec, 4);
	// EfSet: edge, face;  EeSet: edge, edge;  VfSet: vertex, face
	const int lim0 = i < efN + eeN ? ne : nv, lim1 = i < efN ? nf : (i < efN + eeN ? ne : nf);
	const bool valid = id0 >= 0 && id1 >= 0, inside = id0 < lim0 && id1 < lim1;
	if (valid && !inside) atomicOr(inputErr, 1);
	flag[i] = (valid && inside) ? 1 : 0;
}

__global__ void stencil_build_kernel(const unsigned char* __restrict__ ef, const unsigned char* __restrict__ ee,
	const unsigned char* __restrict__ vf, int efN, int eeN, int total, int fix, const int* __restrict__ flag,
	const int* __restrict__ slot, const int4* __restrict__ edges, const int4* __restrict__ faces,
	const int* __restrict__ o2s, Stencil* __restrict__ out, int* __restrict__ outIdx)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total || !flag[i]) return;
	const size_t eeAt = fix ? (size_t)(i - efN) : (size_t)i, vfAt = fix ? (size_t)(i - efN - eeN) : (size_t)i;
	Stencil s;
	s.pad_[0] = s.pad_[1] = s.pad_[2] = 0.f;
	for (int k = 0; k < 5; ++k) { s.index[k] = 0; s.weight[k] = 0.f; }
	if (i < efN)
	{
		const unsigned char* p = ef + 48 * (size_t)i;  // EfSet: eId@0 fId@4 stiff@8 bary@12 normal@32
		int4 e = edges[ldi(p, 0)], f = faces[ldi(p, 4)];
		float b0 = ldf(p, 12), b1 = ldf(p, 16), b2 = ldf(p, 20);
		s.n = 5; s.nFirst = 2;
		s.index[0] = e.x; s.index[1] = e.y; s.index[2] = f.x; s.index[3] = f.y; s.index[4] = f.z;
		s.weight[0] = b0;
		s.weight[1] = __fsub_rn(1.f, b0);
		s.weight[2] = -b1;
		s.weight[3] = -b2;
		s.weight[4] = -__fsub_rn(__fsub_rn(1.f, b1), b2);  // cpp:344-348
	}
	else if (i < efN + eeN)
	{
		const unsigned char* p = ee + 48 * eeAt;  // EeSet: eId0@0 eId1@4 stiff@8 bary@16 normal@32
		int4 e0 = edges[ldi(p, 0)], e1 = edges[ldi(p, 4)];
		float b0 = ldf(p, 16), b1 = ldf(p, 20);
		s.n = 4; s.nFirst = 2;
		s.index[0] = e0.x; s.index[1] = e0.y; s.index[2] = e1.x; s.index[3] = e1.y;
		s.weight[0] = b0;
		s.weight[1] = __fsub_rn(1.f, b0);
		s.weight[2] = -b1;
		s.weight[3] = -__fsub_rn(1.f, b1);  // cpp:372-375
	}
	else
	{
		const unsigned char* p = vf + 48 * vfAt;  // VfSet: vId@0 fId@4 stiff@8 bary@16, Q3: m_bary[2] is the float at byte 24
		int4 f = faces[ldi(p, 4)];
		// fix mode: the third weight is -(1 - b0 - b1), what the padding float stands in for in the literal reading
		float b0 = ldf(p, 16), b1 = ldf(p, 20), b2 = fix ? __fadd_rn(b0, b1) : ldf(p, 24);
		s.n = 4; s.nFirst = 3;
		s.index[0] = f.x; s.index[1] = f.y; s.index[2] = f.z; s.index[3] = ldi(p, 0);
		s.weight[0] = -b0;
		s.weight[1] = -b1;
		s.weight[2] = -__fsub_rn(1.f, b2);
		s.weight[3] = 1.f;  // cpp:397-400
	}
	const unsigned char* rec = i < efN ? ef + 48 * (size_t)i : (i < efN + eeN ? ee + 48 * eeAt : vf + 48 * vfAt);
	s.stiff = ldf(rec, 8);
	for (int k = 0; k < 4; ++k) s.dir[k] = ldf(rec, 32 + 4 * k);
	int dst = slot[i];
	out[dst] = s;
	for (int k = 0; k < 5; ++k) outIdx[5 * dst + k] = k < s.n ? o2s[s.index[k]] : 0;  // cpp:297-300
}

// ------------------------------------------------------------------ scans
// Exclusive scan of `count` ints by a single CTA (counts are per 32-node bank,
// so this is nv/32 elements at most).  total -> *totalOut.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) exclusive_scan_kernel(const int* __restrict__ in, int count,
	int* __restrict__ out, int* __restrict__ totalOut)
{
	__shared__ int warpSum[kScanThreads / 32];
	__shared__ int carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < count; base += kScanThreads)
	{
		int i = base + threadIdx.x;
		int v = i < count ? in[i] : 0;
		int inc = v;
		for (int off = 1; off < 32; off <<= 1)
		{
			int t = __shfl_up_sync(kFull, inc, off);
			if (lane >= off) inc += t;
		}
		if (lane == 31) warpSum[warp] = inc;
		__syncthreads();
		if (warp == 0)
		{
			int w = warpSum[lane];
			int winc = w;
			for (int off = 1; off < 32; off <<= 1)
			{
				int t = __shfl_up_sync(kFull, winc, off);
				if (lane >= off) winc += t;
			}
			warpSum[lane] = winc - w;  // exclusive over warps
		}
		__syncthreads();
		int excl = carry + warpSum[warp] + inc - v;
		if (i < count) out[i] = excl;
		__syncthreads();
		if (threadIdx.x == kScanThreads - 1) carry = excl + v;
		__syncthreads();
	}
	if (threadIdx.x == 0) *totalOut = carry;
}

// ------------------------------------------------------------------ level 0
// BuildConnectMaskL0 (cpp:447-511): bit l of mask[v] set iff lane l of v's bank is v itself or a mesh neighbour.
__global__ void connect_mask_l0_kernel(const int* __restrict__ adjStart, const int* __restrict__ adjIdx, int nv, int nVC,
	unsigned* __restrict__ mask)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nVC) return;
	unsigned m = 0;
	if (v < nv)
	{
		m = 1u << (v & 31);
		int bank = v >> 5;
		for (int e = adjStart[v], end = adjStart[v + 1]; e < end; ++e)
		{
			int u = adjIdx[e];
			if ((u >> 5) == bank) m |= 1u << (u & 31);
		}
	}
	mask[v] = m;
}

// BuildCollisionConnection (cpp:514-563); coarse == nullptr at level 0
__global__ void collision_connect_kernel(const Stencil* __restrict__ st, const int* __restrict__ stIdx, int nStencil,
	const int* __restrict__ coarse, unsigned* __restrict__ mask)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nStencil) return;
	int n = st[i].n, nFirst = st[i].nFirst;
	unsigned id[5], m[5];
	for (int k = 0; k < 5; ++k)
	{
		m[k] = 0;
		int raw = k < n ? stIdx[5 * i + k] : 0;
		id[k] = (unsigned)(coarse && k < n ? coarse[raw] : raw);
	}
	for (int a = 0; a < 5; ++a)
		for (int b = a + 1; b < 5; ++b)
		{
			if (b >= n) continue;
			if (id[a] == id[b]) continue;
			if ((id[a] >> 5) != (id[b] >> 5)) continue;
			if (a < nFirst && b >= nFirst)
			{
				m[a] |= 1u << (id[b] & 31);
				m[b] |= 1u << (id[a] & 31);
			}
		}
	for (int k = 0; k < 5; ++k)
		if (k < n && m[k]) atomicOr(&mask[id[k]], m[k]);
}

// In-bank transitive closure by flood-fill over the 32 masks of a warp
// (cpp:596-614 / 926-944), each lane from its own seed.
__device__ __forceinline__ unsigned close_mask(unsigned own, unsigned lane)
{
	unsigned m = own, seen = 1u << lane;
	while (true)
	{
		unsigned todo = seen ^ m;
		bool more = todo != 0;
		if (!__any_sync(kFull, more)) break;
		int nx = more ? __ffs(todo) - 1 : (int)lane;
		unsigned other = __shfl_sync(kFull, own, nx);
		if (more)
		{
			seen |= 1u << nx;
			m |= other;
		}
	}
	return m;
}

// PreparePrefixSumL0 (cpp:565-628) / NextLevelCluster (cpp:873-961): close the masks of `count` nodes
// (self bit added when addSelf), store them back, and count elected (lowest-lane) nodes per bank.
__global__ void close_components_kernel(unsigned* __restrict__ mask, int count, int addSelf, int* __restrict__ bankCount)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	unsigned lane = threadIdx.x & 31;
	int nBanks = (count + 31) >> 5;
	if (warp >= nBanks) return;
	int node = warp * 32 + lane;
	bool live = node < count;
	unsigned own = live ? mask[node] : 0u;
	if (addSelf || !live) own |= 1u << lane;
	unsigned m = close_mask(own, lane);
	bool elected = live && (m & lanemask_lt(lane)) == 0;
	unsigned ballot = __ballot_sync(kFull, elected);
	if (live) mask[node] = m;
	if (lane == 0) bankCount[warp] = __popc(ballot);
}

// BuildLevel1 (cpp:630-740) / PrefixSumLx (cpp:963-1072): id = #elected in earlier banks + rank of the
// component's lowest lane among this bank's elected lanes.
__global__ void number_components_kernel(const unsigned* __restrict__ mask, int count, const int* __restrict__ bankPrefix,
	int begin, int nextBegin, int* __restrict__ idOut, int* __restrict__ goingNext)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	unsigned lane = threadIdx.x & 31;
	int nBanks = (count + 31) >> 5;
	if (warp >= nBanks) return;
	int node = warp * 32 + lane;
	bool live = node < count;
	unsigned m = live ? mask[node] : (1u << lane);
	bool elected = live && (m & lanemask_lt(lane)) == 0;
	unsigned ballot = __ballot_sync(kFull, elected);
	if (!live) return;
	unsigned rep = __ffs(m) - 1;
	int id = bankPrefix[warp] + __popc(ballot & lanemask_lt(rep));
	idOut[node] = id;
	goingNext[begin + node] = id + nextBegin;
}

// ------------------------------------------------------------------ level >= 1
// BuildConnectMaskLx (cpp:743-871): OR, over every fine vertex of a coarse node, of the bank-local bits of the
// coarse nodes its mesh neighbours belong to.  Lanes that share a coarse node combine before one atomicOr.
__global__ void connect_mask_lx_kernel(const int* __restrict__ adjStart, const int* __restrict__ adjIdx,
	const int* __restrict__ coarse, int nv, unsigned* __restrict__ nextMask)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	bool live = v < nv;
	unsigned cv = live ? (unsigned)coarse[v] : 0xffffffffu;
	unsigned m = 0;
	if (live)
	{
		for (int e = adjStart[v], end = adjStart[v + 1]; e < end; ++e)
		{
			unsigned cu = (unsigned)coarse[adjIdx[e]];
			if ((cu >> 5) == (cv >> 5)) m |= 1u << (cu & 31);
		}
	}
	unsigned peers = __match_any_sync(kFull, cv);
	unsigned combined = __reduce_or_sync(peers, m);
	if (live && combined && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicOr(&nextMask[cv], combined);
}

// ComputeNextLevel (cpp:1074-1084)
__global__ void next_level_table_kernel(const int* __restrict__ coarse, const int* __restrict__ nextId, int nv, int* __restrict__ out)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v < nv) out[v] = nextId[coarse[v]];
}

// AggregationKernel (cpp:1092-1162): ancestors of every vertex at levels 1..numLevel-1
__global__ void coarse_tables_kernel(const int* __restrict__ goingNext, int nv, int numLevel, int4* __restrict__ out)
{
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nv) return;
	int t[4] = { 0, 0, 0, 0 };
	int cur = v;
	for (int l = 0; l < numLevel - 1; ++l)
	{
		cur = goingNext[cur];
		t[l] = cur;
	}
	out[v] = make_int4(t[0], t[1], t[2], t[3]);
}

// Shard cut q+1 (block q): the fine bank nearest to the even split nFine*(q+1)/world, within +-window banks, at which the
// exclusive level-1 id prefix is a multiple of 32 (no level-1 bank straddles the cut).  out[q+1] = bank, out[16+q+1] = 1 if
// aligned, out[32+q+1] = level-1 id at the cut.  No such bank: the even split, flagged unaligned.
__global__ void find_cuts_kernel(const int* __restrict__ bankPrefix, int nBanks, int nFine, int total, int world, int window,
	int* __restrict__ out)
{
	__shared__ unsigned long long best;
	const int q = blockIdx.x + 1;
	const int ideal = (int)((long long)nFine * q / world);
	if (threadIdx.x == 0) best = ~0ull;
	__syncthreads();
	const int lo = max(1, ideal - window), hi = min(nBanks - 1, ideal + window);
	for (int b = lo + (int)threadIdx.x; b <= hi; b += blockDim.x)
		if ((bankPrefix[b] & 31) == 0)
			atomicMin(&best, ((unsigned long long)(unsigned)abs(b - ideal) << 32) | (unsigned)b);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		if (ideal >= nBanks) { out[q] = ideal; out[16 + q] = 1; out[32 + q] = total; }        // nothing but padding behind the cut
		else if (best != ~0ull) { const int b = (int)(best & 0xffffffffu); out[q] = b; out[16 + q] = 1; out[32 + q] = bankPrefix[b]; }
		else { out[q] = ideal; out[16 + q] = 0; out[32 + q] = bankPrefix[ideal]; }
	}
}

}  // namespace

#ifndef MAS_CPU_EMULATION   // host side: launches (tests/emu/cluster_emu.cpp, test infrastructure, has its own launcher)
// Exclusive scan + total.  Up to 64 k elements (meshes up to 2M vertices: one element per 32-node bank) the single CTA above
// is the fastest thing there is (one launch, no temporary storage); beyond that it serialises (151 ms of setup at 33.5M
// vertices went into it) and CUB's decoupled look-back scan takes over, followed by a one-thread kernel for the total.
__global__ void scan_total_kernel(const int* __restrict__ in, const int* __restrict__ out, int count, int* __restrict__ totalOut)
{
	*totalOut = count > 0 ? out[count - 1] + in[count - 1] : 0;
}

int launch_exclusive_scan(Context* c, const int* in, int count, int* out, int* totalOut)
{
	if (count <= 65536)
	{
		exclusive_scan_kernel<<<1, kScanThreads, 0, c->stream>>>(in, count, out, totalOut);
		MAS_CUDA(c, cudaGetLastError());
		return MAS_OK;
	}
	size_t bytes = 0;
	MAS_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, c->stream));
	if (int rc = reserve(c, c->cubTemp, bytes)) return rc;
	MAS_CUDA(c, cub::DeviceScan::ExclusiveSum(c->cubTemp.p, bytes, in, out, count, c->stream));
	scan_total_kernel<<<1, 1, 0, c->stream>>>(in, out, count, totalOut);
	MAS_CUDA(c, cudaGetLastError());
	return MAS_OK;
}

int build_stencils(Context* c, const void* ef, const void* ee, const void* vf, unsigned efN, unsigned eeN, unsigned vfN)
{
	long long total = (long long)efN + eeN + vfN;
	long long cap = (long long)c->nv * kMaxCollisionPerVert;  // cpp:187-188
	if (total > cap) total = cap;                             // cpp:312-316
	c->nStencil = 0;
	if (total <= 0) return MAS_OK;
	cudaStream_t s = c->stream;
	const int n = (int)total, threads = 256;
	if (int rc = reserve(c, c->stencilFlag, (size_t)n)) return rc;
	if (int rc = reserve(c, c->stencilSlot, (size_t)n)) return rc;
	if (int rc = reserve(c, c->scanTotal, 1)) return rc;
	if (int rc = reserve(c, c->inputErr, 1)) return rc;
	MAS_CUDA(c, cudaMemsetAsync(c->inputErr.p, 0, sizeof(int), s));
	stencil_flag_kernel<<<cdiv(n, threads), threads, 0, s>>>((const unsigned char*)ef, (const unsigned char*)ee,
		(const unsigned char*)vf, (int)efN, (int)eeN, n, c->optStencilFix, c->nv, c->ne, c->nf, c->stencilFlag.p, c->inputErr.p);
	if (int rc = launch_exclusive_scan(c, c->stencilFlag.p, n, c->stencilSlot.p, c->scanTotal.p)) return rc;
	c->prepareLaunches += 2;
	int count = 0, bad = 0;
	MAS_CUDA(c, cudaMemcpyAsync(&count, c->scanTotal.p, sizeof(int), cudaMemcpyDeviceToHost, s));
	MAS_CUDA(c, cudaMemcpyAsync(&bad, c->inputErr.p, sizeof(int), cudaMemcpyDeviceToHost, s));
	MAS_CUDA(c, cudaStreamSynchronize(s));
	if (bad)
	{
		c->err = "PreparePreconditioner: a collision stencil names an edge, face or vertex beyond the mesh";
		return MAS_ERR_INVALID;
	}
	c->nStencil = count;
	if (count == 0) return MAS_OK;
	if (int rc = reserve(c, c->stencils, (size_t)count)) return rc;
	if (int rc = reserve(c, c->stencilIdx, (size_t)count * 5)) return rc;
	stencil_build_kernel<<<cdiv(n, threads), threads, 0, s>>>((const unsigned char*)ef, (const unsigned char*)ee,
		(const unsigned char*)vf, (int)efN, (int)eeN, n, c->optStencilFix, c->stencilFlag.p, c->stencilSlot.p, c->edges.p, c->faces.p,
		c->o2s.p, c->stencils.p, c->stencilIdx.p);
	c->prepareLaunches += 1;
	MAS_CUDA(c, cudaGetLastError());
	return MAS_OK;
}

static int number_level(Context* c, unsigned* mask, int count, int addSelf, int begin, int* idOut, int* countOut)
{
	cudaStream_t s = c->stream;
	const int threads = 256;
	const int nBanks = (count + 31) / 32;
	if (int rc = reserve(c, c->bankCount, (size_t)nBanks)) return rc;
	if (int rc = reserve(c, c->bankPrefix, (size_t)nBanks)) return rc;
	if (int rc = reserve(c, c->scanTotal, 1)) return rc;
	close_components_kernel<<<cdiv((long long)nBanks * 32, threads), threads, 0, s>>>(mask, count, addSelf, c->bankCount.p);
	if (int rc = launch_exclusive_scan(c, c->bankCount.p, nBanks, c->bankPrefix.p, c->scanTotal.p)) return rc;
	// node ids of the next level start right after this level's padded range
	const int nextBegin = begin + pad32(count);
	// goingNext must hold [0, nextBegin)
	if ((size_t)nextBegin > c->goingNext.cap)
	{
		TempBuf<int> bigger;
		if (int rc = reserve(c, bigger, (size_t)nextBegin * 2)) return rc;
		MAS_CUDA(c, cudaMemsetAsync(bigger.p, 0, sizeof(int) * bigger.cap, s));
		if (c->goingNext.p) MAS_CUDA(c, cudaMemcpyAsync(bigger.p, c->goingNext.p, sizeof(int) * c->goingNext.cap, cudaMemcpyDeviceToDevice, s));
		MAS_CUDA(c, cudaStreamSynchronize(s));
		std::swap(c->goingNext.p, bigger.p);       // the old array leaves with `bigger`
		std::swap(c->goingNext.cap, bigger.cap);
	}
	number_components_kernel<<<cdiv((long long)nBanks * 32, threads), threads, 0, s>>>(mask, count, c->bankPrefix.p, begin,
		nextBegin, idOut, c->goingNext.p);
	c->prepareLaunches += 3;
	MAS_CUDA(c, cudaMemcpyAsync(countOut, c->scanTotal.p, sizeof(int), cudaMemcpyDeviceToHost, s));
	MAS_CUDA(c, cudaStreamSynchronize(s));
	return MAS_OK;
}

int build_hierarchy(Context* c)
{
	cudaStream_t s = c->stream;
	const int nv = c->nv, nVC = c->nVC, L = c->numLevel, threads = 256;
	for (int l = 0; l <= L; ++l) c->levelSize[l][0] = c->levelSize[l][1] = 0;
	if (int rc = reserve(c, c->fineMask, (size_t)nVC)) return rc;
	for (int l = 0; l < L; ++l)
		if (int rc = reserve(c, c->cst[l], (size_t)nv)) return rc;
	if (c->goingNext.cap < (size_t)nVC + (size_t)nVC / 8 + 4096)
	{
		release(c->goingNext);
		if (int rc = reserve(c, c->goingNext, (size_t)nVC + (size_t)nVC / 8 + 4096)) return rc;
	}
	MAS_CUDA(c, cudaMemsetAsync(c->goingNext.p, 0, sizeof(int) * c->goingNext.cap, s));

	// ---- level 0 -> 1
	connect_mask_l0_kernel<<<cdiv(nVC, threads), threads, 0, s>>>(c->adjStart.p, c->adjIdx.p, nv, nVC, c->fineMask.p);
	c->prepareLaunches += 1;
	if (c->nStencil > 0)
	{
		collision_connect_kernel<<<cdiv(c->nStencil, threads), threads, 0, s>>>(c->stencils.p, c->stencilIdx.p, c->nStencil,
			nullptr, c->fineMask.p);
		c->prepareLaunches += 1;
	}
	int n1 = 0;
	if (int rc = number_level(c, c->fineMask.p, nv, 0, 0, c->cst[0].p, &n1)) return rc;
	c->levelSize[1][0] = n1;
	c->levelSize[1][1] = nVC;
	// level-1 ids are handed out in Morton order of the fine banks, so the level-1 nodes produced by rank q's banks are the
	// contiguous range [bankPrefix[firstBank(q)], bankPrefix[firstBank(q+1)]) — the slice its peers pull from it.  The cuts
	// themselves are chosen here (every rank computes the same ones): see Context::alignedCuts.
	for (int q = 0; q <= 16; ++q) c->l1Slice[q] = n1;
	c->l1Slice[0] = 0;
	c->alignedCuts = false;
	if (c->world > 1)
	{
		const int nBanks = (nv + 31) / 32, nFine = nVC / 32;
		int window = nFine / (8 * c->world);
		if (window < 32) window = 32;
		if (int rc = reserve(c, c->cutInfo, 48)) return rc;
		find_cuts_kernel<<<c->world - 1, 256, 0, s>>>(c->bankPrefix.p, nBanks, nFine, n1, c->world, window, c->cutInfo.p);
		c->prepareLaunches += 1;
		int info[48];
		MAS_CUDA(c, cudaMemcpyAsync(info, c->cutInfo.p, sizeof(info), cudaMemcpyDeviceToHost, s));
		MAS_CUDA(c, cudaStreamSynchronize(s));
		info[0] = 0; info[c->world] = nFine;
		bool aligned = true, ordered = true;
		for (int q = 1; q < c->world; ++q) { aligned = aligned && info[16 + q]; ordered = ordered && info[q] >= info[q - 1] && info[q] <= nFine; }
		if (!ordered || !c->optAlignCuts)
		{
			// windows overlap only on meshes of a few banks per shard (or alignment is switched off): the even split
			aligned = false;
			for (int q = 1; q < c->world; ++q)
			{
				info[q] = (int)((long long)nFine * q / c->world);
				if (info[q] < nBanks) MAS_CUDA(c, cudaMemcpyAsync(&info[32 + q], c->bankPrefix.p + info[q], sizeof(int), cudaMemcpyDeviceToHost, s));
				else info[32 + q] = n1;
			}
			MAS_CUDA(c, cudaStreamSynchronize(s));
		}
		for (int q = 1; q < c->world; ++q) c->l1Slice[q] = info[32 + q];
		c->ownFineBegin = info[c->rank];
		c->ownFineEnd = info[c->rank + 1];
		c->alignedCuts = aligned;
	}
	c->nL1Blocks = pad32(n1) / 32;
	c->l1BlockBegin = 0;
	c->l1BlockEnd = c->nL1Blocks;
	if (c->world > 1)
	{
		const int a = c->l1Slice[c->rank], b = c->l1Slice[c->rank + 1];
		c->l1BlockBegin = a / 32;
		c->l1BlockEnd = b > a ? (b + 31) / 32 : c->l1BlockBegin;
	}

	// ---- level l -> l+1
	for (int level = 1; level < L; ++level)
	{
		const int cnt = c->levelSize[level][0], begin = c->levelSize[level][1];
		if (int rc = reserve(c, c->nextMask, (size_t)pad32(cnt) + 32)) return rc;
		if (int rc = reserve(c, c->nextId, (size_t)pad32(cnt) + 32)) return rc;
		MAS_CUDA(c, cudaMemsetAsync(c->nextMask.p, 0, sizeof(unsigned) * ((size_t)pad32(cnt) + 32), s));
		connect_mask_lx_kernel<<<cdiv(pad32(nv), threads), threads, 0, s>>>(c->adjStart.p, c->adjIdx.p, c->cst[level - 1].p, nv,
			c->nextMask.p);
		c->prepareLaunches += 1;
		if (c->nStencil > 0)
		{
			collision_connect_kernel<<<cdiv(c->nStencil, threads), threads, 0, s>>>(c->stencils.p, c->stencilIdx.p, c->nStencil,
				c->cst[level - 1].p, c->nextMask.p);
			c->prepareLaunches += 1;
		}
		int nNext = 0;
		if (int rc = number_level(c, c->nextMask.p, cnt, 1, begin, c->nextId.p, &nNext)) return rc;
		c->levelSize[level + 1][0] = nNext;
		c->levelSize[level + 1][1] = begin + pad32(cnt);
		if (level == 1 && c->world > 1 && c->alignedCuts)
		{
			// level-2 ids follow the level-1 banks in order: rank q's level-1 banks [l1Slice[q]/32, l1Slice[q+1]/32) produce the
			// level-2 nodes [bankPrefix[l1Slice[q]/32], bankPrefix[l1Slice[q+1]/32))
			const int nL1Banks = (cnt + 31) / 32;
			for (int q = 0; q <= 16; ++q) c->l2Slice[q] = nNext;
			c->l2Slice[0] = 0;
			for (int q = 1; q < c->world; ++q)
			{
				const int bank = c->l1Slice[q] / 32;
				if (bank < nL1Banks) MAS_CUDA(c, cudaMemcpyAsync(&c->l2Slice[q], c->bankPrefix.p + bank, sizeof(int), cudaMemcpyDeviceToHost, s));
			}
			MAS_CUDA(c, cudaStreamSynchronize(s));
		}
		next_level_table_kernel<<<cdiv(nv, threads), threads, 0, s>>>(c->cst[level - 1].p, c->nextId.p, nv, c->cst[level].p);
		c->prepareLaunches += 1;
	}
	c->totalClusters = c->levelSize[L][1];  // cpp:1088
	c->nBlocks = c->totalClusters / 32;
	c->nFineBlocks = nVC / 32;
	c->nCoarseNodes = c->totalClusters - nVC;

	if (int rc = reserve(c, c->coarseTables, (size_t)nv)) return rc;
	coarse_tables_kernel<<<cdiv(nv, threads), threads, 0, s>>>(c->goingNext.p, nv, L, c->coarseTables.p);
	c->prepareLaunches += 1;
	MAS_CUDA(c, cudaGetLastError());
	return MAS_OK;
}

#endif  // MAS_CPU_EMULATION

}  // namespace mas
