// extern "C" boundary of libmas_b200.so (include/mas_b200.h): argument checking, host<->device staging,
// CUDA-graph capture of the apply sequence, and read-only introspection for the parity tests.
#include "mas_internal.h"
#include <cstdlib>

#include <cstdio>
#include <cstring>

namespace mas {

bool check(Context* c, cudaError_t e, const char* what)
{
	if (e == cudaSuccess) return true;
	if (c)
	{
		c->err = std::string(what) + ": " + cudaGetErrorString(e);
	}
	return false;
}

static int fail(Context* c, int code, const char* msg)
{
	if (c) c->err = msg;
	return code;
}

static int level_count(int nv)  // ComputeLevelNums (cpp:112-135): number of levels only; sizes come from real counts
{
	int n = 1, sz = pad32(nv);
	while (sz > 32)
	{
		sz /= 32;
		++n;
		sz = pad32(sz);
	}
	return n;
}

template <typename T>
static int stage_in(Context* c, DevBuf<T>& buf, const void* src, size_t count, int mem, const T** out)
{
	if (mem == MAS_MEM_DEVICE)
	{
		*out = reinterpret_cast<const T*>(src);
		return MAS_OK;
	}
	if (int rc = reserve(c, buf, count)) return rc;
	if (count) MAS_CUDA(c, cudaMemcpyAsync(buf.p, src, sizeof(T) * count, cudaMemcpyHostToDevice, c->stream));
	*out = buf.p;
	return MAS_OK;
}

// MAS_OPT_HOST_PULL: the residual of a host-pointer apply read straight out of page-locked host memory (its device mapping)
// with coalesced 16-byte loads, four independent loads per thread in flight, and stored to the staging buffer in HBM.
__global__ void __launch_bounds__(256) pull_host_kernel(const float4* __restrict__ mapped, float4* __restrict__ dst, int n)
{
	const int stride = gridDim.x * blockDim.x;
	for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride)
	{
		float4 v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (i0 + u * stride < n) v[u] = mapped[i0 + u * stride];
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (i0 + u * stride < n) dst[i0 + u * stride] = v[u];
	}
}

// Sharded host-pointer apply: only the vertices this rank owns cross PCIe.  Owned vertices are a contiguous range in sorted
// (Morton) order and scattered in the caller's order; consecutive sorted vertices are spatial neighbours, so their original
// indices come in short runs and the 16-byte accesses of a warp merge into a few PCIe requests.  Four vertices per thread
// in flight.  gather: dst[ov] = src[ov] (host -> staging);  the same kernel with the roles swapped returns z.
__global__ void __launch_bounds__(256) copy_owned_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
	const int* __restrict__ s2o, int vBegin, int vEnd)
{
	const int stride = gridDim.x * blockDim.x;
	for (int v0 = vBegin + blockIdx.x * blockDim.x + threadIdx.x; v0 < vEnd; v0 += 4 * stride)
	{
		int ov[4];
		float4 val[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) ov[u] = v0 + u * stride < vEnd ? s2o[v0 + u * stride] : -1;
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (ov[u] >= 0) val[u] = src[ov[u]];
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (ov[u] >= 0) dst[ov[u]] = val[u];
	}
}

// device-side address of a page-locked host buffer, or nullptr for pageable memory (which only the copy engine can read)
static const float4* mapped_host_pointer(const void* host)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, host) != cudaSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
	return reinterpret_cast<const float4*>(at.devicePointer);
}

// MAS_OPT_REGISTER_HOST: page-lock a pageable caller buffer where it lies, once, so that later copies run at pinned speed.
// At most four ranges per context (r and z of a solver, twice); the oldest is unlocked when a fifth shows up.
static void unregister_host_ranges(Context* c)
{
	for (auto& r : c->registered)
	{
		if (r.p) cudaHostUnregister(const_cast<void*>(r.p));
		r.p = nullptr;
		r.bytes = 0;
	}
	cudaGetLastError();
}

static void register_host_range(Context* c, const void* p, size_t bytes)
{
	for (const auto& r : c->registered)
		if (r.p == p && r.bytes >= bytes) return;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
	if (at.type != cudaMemoryTypeUnregistered) return;                 // already page-locked (or not host memory at all)
	if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return; }
	Context::HostRange* slot = nullptr;
	for (auto& r : c->registered)
		if (!r.p) { slot = &r; break; }
	if (!slot)
	{
		cudaHostUnregister(const_cast<void*>(c->registered[0].p));
		for (int k = 0; k + 1 < 4; ++k) c->registered[k] = c->registered[k + 1];
		slot = &c->registered[3];
	}
	slot->p = p;
	slot->bytes = bytes;
}

static void drop_graph(Context* c)
{
	for (Context::ApplyGraphSlot& g : c->applyGraphs)
	{
		if (g.exec) cudaGraphExecDestroy(g.exec);
		g = Context::ApplyGraphSlot();
	}
}

static void close_peers(Context* c)
{
	for (int q = 0; q < 16; ++q)
	{
		if (c->peerOpened[q] && c->peerArena[q]) cudaIpcCloseMemHandle(c->peerArena[q]);
		c->peerOpened[q] = false;
		c->peerArena[q] = nullptr;
	}
	c->p2p = false;
}

// sticky: a peer wait of the sharded apply timed out at some point (a rank died or never launched); z is not to be trusted
static int peer_failed(Context* c)
{
	const unsigned code = c->peerErrHost ? *reinterpret_cast<volatile unsigned*>(c->peerErrHost) : 0u;
	if (code == 1u)
	{
		c->err = "peer-memory exchange timed out: a rank stopped publishing its coarse residuals (results are invalid)";
		return MAS_ERR_CUDA;
	}
	if (code != 0u)
	{
		c->err = "apply: a device-side wait timed out (results are invalid)";
		return MAS_ERR_CUDA;
	}
	return MAS_OK;
}

static void free_all(Context* c)
{
	drop_graph(c);
	close_peers(c);
	if (c->peerErrHost) { cudaFreeHost(c->peerErrHost); c->peerErrHost = nullptr; c->peerErrDev = nullptr; }
	unregister_host_ranges(c);
	release(c->arena); release(c->cutInfo);
	release(c->positions); release(c->edges); release(c->faces); release(c->inStarts); release(c->inIdx);
	release(c->aabb); release(c->code); release(c->codeSorted); release(c->s2o); release(c->o2s); release(c->iota);
	release(c->adjStart); release(c->adjIdx); release(c->cubTemp);
	release(c->stencils); release(c->stencilIdx); release(c->stencilFlag); release(c->stencilSlot);
	release(c->fineMask);
	for (int l = 0; l < kMaxLevel; ++l) release(c->cst[l]);
	release(c->goingNext); release(c->nextMask); release(c->nextId); release(c->bankCount); release(c->bankPrefix);
	release(c->scanTotal); release(c->coarseTables);
	release(c->diagIn); release(c->offdiagIn); release(c->rangesIn); release(c->efIn); release(c->eeIn); release(c->vfIn);
	release(c->extraFine); release(c->cooCount); release(c->cooStart); release(c->cooFill); release(c->cooVal);
	release(c->coarseAcc); release(c->packedInv); release(c->posTab); release(c->posTab96); release(c->invertErr); release(c->inputErr);
	release(c->coarseR); release(c->coarseZ); release(c->coarseZsum); release(c->rIn); release(c->zOut);
	release(c->pcgR); release(c->pcgB); release(c->pcgX);
	release(c->pcgPartials); release(c->pcgState); release(c->pcgDiag); release(c->pcgOff); release(c->pcgRanges); release(c->pcgIdx);
	release(c->pcgSliceSlots); release(c->pcgSliceStart); release(c->pcgEllIdx); release(c->pcgEllVal);
}

}  // namespace mas

using namespace mas;

extern "C" {

int mas_create(mas_handle_t* out, int device)
{
	if (!out) return MAS_ERR_INVALID;
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return MAS_ERR_CUDA;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MAS_ERR_CUDA;
	if (prop.major != 10) return MAS_ERR_CUDA;  // sm_100a code only; no fallback path exists
	if (cudaSetDevice(device) != cudaSuccess) return MAS_ERR_CUDA;
	mas_context* c = new mas_context();
	c->device = device;
	c->smCount = prop.multiProcessorCount;
	cudaEventCreate(&c->evA);
	cudaEventCreate(&c->evB);
	cudaEventCreate(&c->evAp0);
	cudaEventCreate(&c->evAp1);
	cudaEventCreate(&c->evF0);
	cudaEventCreate(&c->evF1);
	cudaEventCreate(&c->evS0);
	cudaEventCreate(&c->evS1);
	cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&c->evHead, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&c->evCoarse, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&c->evTail, cudaEventDisableTiming);
	{
		int prLow = 0, prHigh = 0;
		cudaDeviceGetStreamPriorityRange(&prLow, &prHigh);
		cudaStreamCreateWithPriority(&c->sideA, cudaStreamNonBlocking, prLow);
		cudaStreamCreateWithPriority(&c->sideB, cudaStreamNonBlocking, prLow);
	}
	// sticky error word of the apply kernels in page-locked, device-mapped host memory (the host sees a timed-out device-side
	// peer wait without a device round trip)
	if (cudaHostAlloc((void**)&c->peerErrHost, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
		cudaHostGetDevicePointer((void**)&c->peerErrDev, c->peerErrHost, 0) != cudaSuccess)
	{
		delete c;
		return MAS_ERR_CUDA;
	}
	*c->peerErrHost = 0u;
	*out = c;
	return MAS_OK;
}

int mas_destroy(mas_handle_t h)
{
	if (!h) return MAS_ERR_INVALID;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	free_all(h);
	if (h->evA) cudaEventDestroy(h->evA);
	if (h->evB) cudaEventDestroy(h->evB);
	if (h->evAp0) cudaEventDestroy(h->evAp0);
	if (h->evAp1) cudaEventDestroy(h->evAp1);
	if (h->evF0) cudaEventDestroy(h->evF0);
	if (h->evF1) cudaEventDestroy(h->evF1);
	if (h->evS0) cudaEventDestroy(h->evS0);
	if (h->evS1) cudaEventDestroy(h->evS1);
	if (h->evFork) cudaEventDestroy(h->evFork);
	if (h->evHead) cudaEventDestroy(h->evHead);
	if (h->evCoarse) cudaEventDestroy(h->evCoarse);
	if (h->evTail) cudaEventDestroy(h->evTail);
	if (h->sideA) cudaStreamDestroy(h->sideA);
	if (h->sideB) cudaStreamDestroy(h->sideB);
	delete h;
	return MAS_OK;
}

const char* mas_last_error(mas_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int mas_set_stream(mas_handle_t h, void* cuda_stream)
{
	if (!h) return MAS_ERR_INVALID;
	h->stream = (cudaStream_t)cuda_stream;
	drop_graph(h);
	return MAS_OK;
}

int mas_set_option(mas_handle_t h, int key, int value)
{
	if (!h) return MAS_ERR_INVALID;
	switch (key)
	{
	case MAS_OPT_PROLONG_ALL_LEVELS: h->optProlongAll = value ? 1 : 0; break;
	case MAS_OPT_APPLY_VARIANT: h->optApplyVariant = value; break;
	case MAS_OPT_USE_GRAPH: h->optUseGraph = value ? 1 : 0; break;
	case MAS_OPT_TIME_KERNELS: h->optTimeKernels = value ? 1 : 0; break;
	case MAS_OPT_ALIGN_CUTS: h->optAlignCuts = value ? 1 : 0; h->hierarchyCached = false; break;
	case MAS_OPT_CACHE_HIERARCHY: h->optCacheHierarchy = value ? 1 : 0; return MAS_OK;
	case MAS_OPT_STRICT_PUBLISH: h->optStrictPublish = value ? 1 : 0; break;
	case MAS_OPT_PCG_PERSIST_L2: h->optPcgPersistL2 = value ? 1 : 0; break;
	case MAS_OPT_STENCIL_FIX: h->optStencilFix = value ? 1 : 0; break;
	case MAS_OPT_RESORT_PERIOD: h->optResortPeriod = value > 0 ? value : 0; break;
	case MAS_OPT_INVERT_VARIANT:
		if (value < 0 || value > 1) return fail(h, MAS_ERR_INVALID, "MAS_OPT_INVERT_VARIANT takes 0 (tensor cores) or 1 (FP32 CUDA cores)");
		h->optInvertVariant = value;
		return MAS_OK;   // takes effect at the next mas_prepare
	case MAS_OPT_REGISTER_HOST:
		h->optRegisterHost = value ? 1 : 0;
		if (!value) unregister_host_ranges(h);
		return MAS_OK;
	case MAS_OPT_HOST_PULL:   // staging only: the apply graph stays valid
		if (value < 0 || value > 2) return fail(h, MAS_ERR_INVALID, "MAS_OPT_HOST_PULL takes 0, 1 or 2");
		h->optHostPull = value;
		h->pullCalls = 0; h->pullChoice = -1; h->pullBestMs[0] = h->pullBestMs[1] = 1e30f;
		return MAS_OK;
	default: return fail(h, MAS_ERR_INVALID, "unknown option");
	}
	drop_graph(h);
	return MAS_OK;
}

int mas_set_partition(mas_handle_t h, int rank, int world)
{
	if (!h || world < 1 || rank < 0 || rank >= world) return MAS_ERR_INVALID;
	if (h->allocated) return fail(h, MAS_ERR_INVALID, "mas_set_partition must precede mas_allocate");
	if (world > 16) return fail(h, MAS_ERR_UNSUPPORTED, "at most 16 ranks (the GPUs of one box)");
	h->rank = rank;
	h->world = world;
	return MAS_OK;
}

// Input check of mas_allocate, before anything dereferences the caller's indices: the adjacency rows are ordered
// (starts[0] = 0, non-decreasing, starts[nv] = nnz) and every neighbour, edge and face vertex lies inside the mesh.  The
// reference trusts its caller and reads out of bounds otherwise.
__global__ void validate_mesh_kernel(const int* __restrict__ starts, const int* __restrict__ idx, const int4* __restrict__ edges,
	const int4* __restrict__ faces, int nv, int nnz, int ne, int nf, int* __restrict__ inputErr)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int bad = 0;
	if (i < nv && (starts[i] > starts[i + 1] || (i == 0 && starts[0] != 0) || (i == nv - 1 && starts[nv] != nnz))) bad |= 1;
	if (i < nnz && (idx[i] < 0 || idx[i] >= nv)) bad |= 2;
	if (i < ne)
	{
		const int4 e = edges[i];
		if (e.x < 0 || e.x >= nv || e.y < 0 || e.y >= nv) bad |= 4;
	}
	if (i < nf)
	{
		const int4 f = faces[i];
		if (f.x < 0 || f.x >= nv || f.y < 0 || f.y >= nv || f.z < 0 || f.z >= nv) bad |= 8;
	}
	if (bad) atomicOr(inputErr, bad);
}

int mas_allocate(mas_handle_t h, int numVerts, int numEdges, int numFaces, const float* positions, const int* edges,
	const int* faces, const int* nbrStarts, const int* nbrIdx, int mem)
{
	if (!h || numVerts <= 0 || !positions || !nbrStarts) return MAS_ERR_INVALID;
	Context* c = h;
	MAS_CUDA(c, cudaSetDevice(c->device));
	// Q1 (cpp:44-64): m_frameIndex sticks at 1 after the first call, so the reference sorts exactly once per object;
	// MAS_OPT_RESORT_PERIOD > 0 rebuilds the order every that many calls (same sizes)
	c->allocateCalls += 1;
	if (c->allocated)
	{
		// The reference would carry on with the buffers and the order of the first mesh (and run out of bounds); a different
		// mesh needs a new object.
		if (numVerts != c->nv || numEdges != c->ne || numFaces != c->nf)
			return fail(c, MAS_ERR_INVALID, "AllocatePrecoditioner: this object was allocated for a mesh of other sizes (one mesh per object, cpp:44-64)");
		const bool resort = c->optResortPeriod > 0 && (c->allocateCalls - 1) % c->optResortPeriod == 0;
		if (!resort) return MAS_OK;
		drop_graph(c);
	}
	c->nv = numVerts; c->ne = numEdges; c->nf = numFaces;
	c->nVC = pad32(numVerts);
	c->numLevel = level_count(numVerts);
	if (c->numLevel > kMaxLevel) return fail(c, MAS_ERR_UNSUPPORTED, "more than 5 levels (numVerts > 33,554,432)");

	int nnz = 0;
	if (mem == MAS_MEM_DEVICE)
	{
		MAS_CUDA(c, cudaMemcpyAsync(&nnz, nbrStarts + numVerts, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	else
		nnz = nbrStarts[numVerts];
	if (nnz < 0 || (nnz > 0 && !nbrIdx)) return fail(c, MAS_ERR_INVALID, "bad adjacency");
	c->nnz = nnz;

	const float4* dPos; const int* dStarts; const int* dIdx; const int4* dEdges; const int4* dFaces;
	if (int rc = stage_in(c, c->positions, positions, (size_t)numVerts, mem, &dPos)) return rc;
	if (int rc = stage_in(c, c->inStarts, nbrStarts, (size_t)numVerts + 1, mem, &dStarts)) return rc;
	if (int rc = stage_in(c, c->inIdx, nbrIdx, (size_t)nnz, mem, &dIdx)) return rc;
	// edges/faces are dereferenced again in every PreparePreconditioner (cpp:332-333): keep a device copy
	if (int rc = reserve(c, c->edges, (size_t)(numEdges > 0 ? numEdges : 1))) return rc;
	if (int rc = reserve(c, c->faces, (size_t)(numFaces > 0 ? numFaces : 1))) return rc;
	cudaMemcpyKind kind = mem == MAS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	if (numEdges > 0 && edges) MAS_CUDA(c, cudaMemcpyAsync(c->edges.p, edges, sizeof(int4) * (size_t)numEdges, kind, c->stream));
	if (numFaces > 0 && faces) MAS_CUDA(c, cudaMemcpyAsync(c->faces.p, faces, sizeof(int4) * (size_t)numFaces, kind, c->stream));
	(void)dEdges; (void)dFaces;
	{
		if (int rc = reserve(c, c->inputErr, 1)) return rc;
		MAS_CUDA(c, cudaMemsetAsync(c->inputErr.p, 0, sizeof(int), c->stream));
		long long most = numVerts;
		for (long long n : { (long long)nnz, (long long)(edges ? numEdges : 0), (long long)(faces ? numFaces : 0) }) most = n > most ? n : most;
		validate_mesh_kernel<<<(unsigned)cdiv(most, 256), 256, 0, c->stream>>>(dStarts, dIdx, c->edges.p, c->faces.p, numVerts, nnz,
			edges ? numEdges : 0, faces ? numFaces : 0, c->inputErr.p);
		int bad = 0;
		MAS_CUDA(c, cudaMemcpyAsync(&bad, c->inputErr.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		MAS_CUDA(c, cudaStreamSynchronize(c->stream));
		if (bad)
			return fail(c, MAS_ERR_INVALID, bad & 1 ? "AllocatePrecoditioner: the adjacency row starts are not ordered from 0 to nnz"
				: bad & 2 ? "AllocatePrecoditioner: a neighbour index lies outside the mesh"
				: bad & 4 ? "AllocatePrecoditioner: an edge names a vertex outside the mesh"
				: "AllocatePrecoditioner: a face names a vertex outside the mesh");
	}

	// Morton-contiguous shard of fine banks for this rank
	const int nFine = c->nVC / 32;
	c->ownFineBegin = (int)((long long)nFine * c->rank / c->world);
	c->ownFineEnd = (int)((long long)nFine * (c->rank + 1) / c->world);

	if (c->world > 1 && !c->arena.p)   // (kept across re-sorts: the peers have it mapped)
	{
		// peer-writable arena: two receive buffers for the coarse residuals + control words.  Coarse nodes number about
		// nv/31 on meshes whose banks stay connected; 1/8 of the vertices (+ slack) is the capacity, checked in mas_prepare.
		c->arenaCap = (size_t)c->nVC / 8 + 8192;
		const size_t bytes = 2 * sizeof(float4) * c->arenaCap + 256;
		if (int rc = reserve(c, c->arena, bytes)) return rc;
		MAS_CUDA(c, cudaMemsetAsync(c->arena.p, 0, bytes, c->stream));
		MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	c->hierarchyCached = false;
	if (int rc = order_vertices(c, dPos, dStarts, dIdx)) return rc;
	c->allocated = true;
	c->prepared = false;
	return MAS_OK;
}

int mas_prepare_begin(mas_handle_t h, const float* diagonal, const float* csrOffDiagonals, const int* csrRanges,
	const void* efSets, const void* eeSets, const void* vfSets, unsigned efTotal, unsigned eeTotal, unsigned vfTotal, int mem)
{
	if (!h || !diagonal || !csrRanges) return MAS_ERR_INVALID;
	Context* c = h;
	if (!c->allocated) return fail(c, MAS_ERR_INVALID, "mas_allocate first");
	MAS_CUDA(c, cudaSetDevice(c->device));
	drop_graph(c);
	c->prepared = false;
	c->prepareLaunches = 0;
	MAS_CUDA(c, cudaEventRecord(c->evA, c->stream));

	const float *dDiag, *dOff; const int* dRanges;
	if (int rc = stage_in(c, c->diagIn, diagonal, (size_t)c->nv * 9, mem, &dDiag)) return rc;
	if (int rc = stage_in(c, c->offdiagIn, csrOffDiagonals, (size_t)c->nnz * 9, mem, &dOff)) return rc;
	if (int rc = stage_in(c, c->rangesIn, csrRanges, (size_t)c->nv + 1, mem, &dRanges)) return rc;
	const unsigned long long total = (unsigned long long)efTotal + eeTotal + vfTotal;
	const unsigned char *dEf = nullptr, *dEe = nullptr, *dVf = nullptr;
	if (total > 0)
	{
		if (!efSets || !eeSets || !vfSets) return fail(c, MAS_ERR_INVALID, "stencil totals > 0 but a set pointer is null");
		// Q2: each kind is read at the GLOBAL stencil index, so the caller's arrays span that far (cpp:328/357/383)
		size_t efBytes = 48 * (size_t)efTotal, eeBytes = 48 * (size_t)(efTotal + eeTotal), vfBytes = 48 * (size_t)total;
		if (c->optStencilFix) { eeBytes = 48 * (size_t)eeTotal; vfBytes = 48 * (size_t)vfTotal; }
		if (int rc = stage_in(c, c->efIn, efSets, efBytes, mem, &dEf)) return rc;
		if (int rc = stage_in(c, c->eeIn, eeSets, eeBytes, mem, &dEe)) return rc;
		if (int rc = stage_in(c, c->vfIn, vfSets, vfBytes, mem, &dVf)) return rc;
	}
	if (int rc = build_stencils(c, dEf, dEe, dVf, efTotal, eeTotal, vfTotal)) return rc;
	// The clustering depends on the sorted adjacency (fixed since mas_allocate) and on the stencils only: without stencils
	// two consecutive prepares build the same hierarchy, bit for bit (MAS_OPT_CACHE_HIERARCHY skips the second build).
	if (!(c->optCacheHierarchy && c->hierarchyCached && c->nStencil == 0))
	{
		c->hierarchyCached = false;
		if (int rc = build_hierarchy(c)) return rc;
		c->hierarchyCached = c->nStencil == 0;
	}

	if (c->p2p && (size_t)c->nCoarseNodes > c->arenaCap)
		return fail(c, MAS_ERR_UNSUPPORTED, "coarse hierarchy larger than the peer arena: use the begin/exchange/end protocol");
	// apply-side buffers, sized from the actual hierarchy; padding slots stay zero for the lifetime of this setup
	const size_t nc = (size_t)(c->nCoarseNodes > 0 ? c->nCoarseNodes : 1);
	if (int rc = reserve(c, c->coarseR, nc)) return rc;
	if (int rc = reserve(c, c->coarseZ, nc)) return rc;
	if (int rc = reserve(c, c->coarseZsum, nc)) return rc;
	MAS_CUDA(c, cudaMemsetAsync(c->coarseR.p, 0, sizeof(float4) * nc, c->stream));
	MAS_CUDA(c, cudaMemsetAsync(c->coarseZ.p, 0, sizeof(float4) * nc, c->stream));
	MAS_CUDA(c, cudaMemsetAsync(c->coarseZsum.p, 0, sizeof(float4) * nc, c->stream));

	return assemble_and_invert_begin(c, dDiag, dOff, dRanges);
}

int mas_prepare_end(mas_handle_t h)
{
	if (!h) return MAS_ERR_INVALID;
	Context* c = h;
	MAS_CUDA(c, cudaSetDevice(c->device));
	if (int rc = assemble_and_invert_end(c)) return rc;
	MAS_CUDA(c, cudaEventRecord(c->evB, c->stream));
	int invertErr = 0;
	if (c->invertErr.p) MAS_CUDA(c, cudaMemcpyAsync(&invertErr, c->invertErr.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaEventElapsedTime(&c->lastPrepareMs, c->evA, c->evB);
	if (invertErr) return fail(c, MAS_ERR_CUDA, "tensor-core inversion: an MMA completion wait timed out (inverses are invalid)");
	c->prepared = true;
	return MAS_OK;
}

int mas_prepare(mas_handle_t h, const float* diagonal, const float* csrOffDiagonals, const int* csrRanges, const void* efSets,
	const void* eeSets, const void* vfSets, unsigned efTotal, unsigned eeTotal, unsigned vfTotal, int mem)
{
	if (int rc = mas_prepare_begin(h, diagonal, csrOffDiagonals, csrRanges, efSets, eeSets, vfSets, efTotal, eeTotal, vfTotal, mem)) return rc;
	return mas_prepare_end(h);
}

static int run_apply_device(Context* c, const float4* r, float4* z)
{
	if ((c->world > 1 && !c->p2p) || !c->optUseGraph || c->optTimeKernels)
	{
		c->applyLaunches = 0;
		if (c->optTimeKernels) MAS_CUDA(c, cudaEventRecord(c->evAp0, c->stream));
		if (int rc = apply_begin(c, r)) return rc;
		if (int rc = apply_end(c, r, z)) return rc;
		if (c->optTimeKernels) MAS_CUDA(c, cudaEventRecord(c->evAp1, c->stream));
		return MAS_OK;
	}
	Context::ApplyGraphSlot* slot = nullptr;
	for (Context::ApplyGraphSlot& g : c->applyGraphs)
		if (g.exec && g.r == (const float*)r && g.z == (float*)z) slot = &g;
	if (!slot)
	{
		slot = &c->applyGraphs[0];
		for (Context::ApplyGraphSlot& g : c->applyGraphs)
			if (!g.exec || (slot->exec && g.lastUse < slot->lastUse)) slot = &g;
		if (slot->exec) cudaGraphExecDestroy(slot->exec);
		*slot = Context::ApplyGraphSlot();
		// the capture origin carries the latency-bound coarse chain: highest priority, so that its CTAs are dispatched ahead of
		// the queued fine-level CTAs of the concurrent branch (kernel nodes inherit the capturing stream's priority)
		cudaStream_t cap;
		int prLow = 0, prHigh = 0;
		MAS_CUDA(c, cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
		MAS_CUDA(c, cudaStreamCreateWithPriority(&cap, cudaStreamNonBlocking, prHigh));
		cudaStream_t saved = c->stream;
		c->stream = cap;
		c->applyLaunches = 0;
		cudaGraph_t graph = nullptr;
		int rc = MAS_OK;
		if (!check(c, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) rc = MAS_ERR_CUDA;
		if (rc == MAS_OK) rc = apply_forked(c, r, z, cap);
		cudaError_t e = cudaStreamEndCapture(cap, &graph);
		c->stream = saved;
		if (rc == MAS_OK && !check(c, e, "cudaStreamEndCapture")) rc = MAS_ERR_CUDA;
		if (rc == MAS_OK) rc = prioritize_apply_graph(c, graph);
		if (rc == MAS_OK && !check(c, cudaGraphInstantiate(&slot->exec, graph, 0), "cudaGraphInstantiate")) rc = MAS_ERR_CUDA;
		if (graph) cudaGraphDestroy(graph);
		cudaStreamDestroy(cap);
		if (rc != MAS_OK) { slot->exec = nullptr; return rc; }
		slot->r = (const float*)r;
		slot->z = (float*)z;
	}
	slot->lastUse = ++c->applyGraphClock;
	MAS_CUDA(c, cudaGraphLaunch(slot->exec, c->stream));
	return MAS_OK;
}

int mas_apply(mas_handle_t h, float* z, const float* residual, int mem)
{
	if (!h || !z || !residual) return MAS_ERR_INVALID;
	Context* c = h;
	if (!c->prepared) return fail(c, MAS_ERR_INVALID, "mas_prepare first");
	if (c->world > 1 && !c->p2p)
		return fail(c, MAS_ERR_INVALID, "sharded context: attach the peers (mas_peer_attach) or use mas_apply_begin / exchange / mas_apply_end");
	MAS_CUDA(c, cudaSetDevice(c->device));
	if (int rc = peer_failed(c)) return rc;        // sticky: an earlier apply lost a peer
	if (mem == MAS_MEM_DEVICE)
	{
		// the level-0 solve writes z while other kernels of the same graph still read r (host pointers are staged through
		// separate device buffers, so an in-place call works there as it does in the reference)
		const char* zb = (const char*)z; const char* rb = (const char*)residual;
		const size_t bytes = sizeof(float4) * (size_t)c->nv;
		if (zb < rb + bytes && rb < zb + bytes) return fail(c, MAS_ERR_INVALID, "mas_apply: z and residual overlap in device memory");
		return run_apply_device(c, (const float4*)residual, (float4*)z);
	}
	if (int rc = reserve(c, c->rIn, (size_t)c->nv)) return rc;
	if (int rc = reserve(c, c->zOut, (size_t)c->nv)) return rc;
	if (c->optRegisterHost)
	{
		register_host_range(c, residual, sizeof(float4) * (size_t)c->nv);
		register_host_range(c, z, sizeof(float4) * (size_t)c->nv);
	}
	const size_t whole = sizeof(float4) * (size_t)c->nv;
	if (c->world > 1)
	{
		// A shard reads r and writes z for ITS OWN vertices only; the caller's z entries of other ranks' vertices stay as they
		// are.  Page-locked buffers (cudaHostAlloc / cudaHostRegister / MAS_OPT_REGISTER_HOST): kernels move just the owned
		// entries through the buffers' device mappings, 2 x 16 B per owned vertex over PCIe.  Pageable buffers can only be
		// reached by the copy engine: whole r in, whole z in (to preserve the foreign entries), whole z out.
		const int vBegin = c->ownFineBegin * 32 < c->nv ? c->ownFineBegin * 32 : c->nv;
		const int vEnd = c->ownFineEnd * 32 < c->nv ? c->ownFineEnd * 32 : c->nv;
		const size_t owned = sizeof(float4) * (size_t)(vEnd - vBegin);
		const float4* mr = mapped_host_pointer(residual);
		float4* mz = const_cast<float4*>(mapped_host_pointer(z));
		int grid = cdiv(vEnd - vBegin, 4 * 256);
		if (grid > 32 * c->smCount) grid = 32 * c->smCount;
		if (grid < 1) grid = 1;
		if (mr) copy_owned_kernel<<<grid, 256, 0, c->stream>>>(mr, c->rIn.p, c->s2o.p, vBegin, vEnd);
		else MAS_CUDA(c, cudaMemcpyAsync(c->rIn.p, residual, whole, cudaMemcpyHostToDevice, c->stream));
		if (!mz) MAS_CUDA(c, cudaMemcpyAsync(c->zOut.p, z, whole, cudaMemcpyHostToDevice, c->stream));
		if (int rc = run_apply_device(c, c->rIn.p, c->zOut.p)) return rc;
		if (mz) copy_owned_kernel<<<grid, 256, 0, c->stream>>>(c->zOut.p, mz, c->s2o.p, vBegin, vEnd);
		else MAS_CUDA(c, cudaMemcpyAsync(z, c->zOut.p, whole, cudaMemcpyDeviceToHost, c->stream));
		MAS_CUDA(c, cudaGetLastError());
		MAS_CUDA(c, cudaStreamSynchronize(c->stream));
		c->hostBytesIn = (long long)((mr ? owned : whole) + (mz ? 0 : whole));
		c->hostBytesOut = (long long)(mz ? owned : whole);
		return peer_failed(c);
	}
	c->hostBytesIn = c->hostBytesOut = (long long)whole;
	const float4* mapped = c->optHostPull ? mapped_host_pointer(residual) : nullptr;
	// auto mode (2): the first six applies with a page-locked residual time the two stagings (three each, the first of each
	// is warm-up) with CUDA events on the stream; the faster one is kept for the rest of the context's life
	const bool sampling = mapped && c->optHostPull == 2 && c->pullChoice < 0;
	const int samplePull = sampling ? (c->pullCalls >= 3 ? 1 : 0) : 0;
	const bool pull = mapped && (c->optHostPull == 1 || (sampling ? samplePull == 1 : c->pullChoice == 1));
	if (sampling) MAS_CUDA(c, cudaEventRecord(c->evS0, c->stream));
	if (pull)
	{
		int grid = cdiv(c->nv, 4 * 256);
		if (grid > 32 * c->smCount) grid = 32 * c->smCount;
		pull_host_kernel<<<grid, 256, 0, c->stream>>>(mapped, c->rIn.p, c->nv);
		MAS_CUDA(c, cudaGetLastError());
	}
	else
		MAS_CUDA(c, cudaMemcpyAsync(c->rIn.p, residual, sizeof(float4) * (size_t)c->nv, cudaMemcpyHostToDevice, c->stream));
	if (sampling) MAS_CUDA(c, cudaEventRecord(c->evS1, c->stream));
	if (int rc = run_apply_device(c, c->rIn.p, c->zOut.p)) return rc;
	MAS_CUDA(c, cudaMemcpyAsync(z, c->zOut.p, sizeof(float4) * (size_t)c->nv, cudaMemcpyDeviceToHost, c->stream));
	MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	if (sampling)
	{
		float ms = 0.f;
		MAS_CUDA(c, cudaEventElapsedTime(&ms, c->evS0, c->evS1));
		if (c->pullCalls % 3 != 0 && ms < c->pullBestMs[samplePull]) c->pullBestMs[samplePull] = ms;
		if (++c->pullCalls >= 6) c->pullChoice = c->pullBestMs[1] < c->pullBestMs[0] ? 1 : 0;
	}
	return MAS_OK;
}

int mas_apply_begin(mas_handle_t h, const float* residual, int mem)
{
	if (!h || !residual) return MAS_ERR_INVALID;
	Context* c = h;
	if (!c->prepared) return fail(c, MAS_ERR_INVALID, "mas_prepare first");
	if (mem != MAS_MEM_DEVICE) return fail(c, MAS_ERR_INVALID, "phase-split apply takes device pointers");
	MAS_CUDA(c, cudaSetDevice(c->device));
	c->applyLaunches = 0;
	c->phaseR = residual;
	// the caller sums the exchange buffer between _begin and _end: attached peers play no part in this protocol
	c->phaseSplit = true;
	const int rc = apply_begin(c, (const float4*)residual);
	c->phaseSplit = false;
	return rc;
}

int mas_apply_end(mas_handle_t h, float* z, int mem)
{
	if (!h || !z) return MAS_ERR_INVALID;
	Context* c = h;
	if (mem != MAS_MEM_DEVICE) return fail(c, MAS_ERR_INVALID, "phase-split apply takes device pointers");
	{
		const char* zb = (const char*)z; const char* rb = (const char*)c->phaseR;
		const size_t bytes = sizeof(float4) * (size_t)c->nv;
		if (rb && zb < rb + bytes && rb < zb + bytes) return fail(c, MAS_ERR_INVALID, "mas_apply_end: z and residual overlap in device memory");
	}
	if (!c->phaseR) return fail(c, MAS_ERR_INVALID, "mas_apply_begin first");
	MAS_CUDA(c, cudaSetDevice(c->device));
	c->phaseSplit = true;
	const int rc = apply_end(c, (const float4*)c->phaseR, (float4*)z);
	c->phaseSplit = false;
	return rc;
}

int mas_pcg_solve(mas_handle_t h, const float* diagonal, const float* csrOffDiagonals, const int* csrRanges, const int* csrIdx,
	const float* b, float* x, float relTol, int maxIter, int usePreconditioner, int mem, int* itersOut, float* relResOut)
{
	if (!h || !diagonal || !csrRanges || !b || !x || maxIter < 0 || !(relTol > 0.f)) return MAS_ERR_INVALID;
	Context* c = h;
	if (!c->allocated) return fail(c, MAS_ERR_INVALID, "mas_allocate first");
	if (usePreconditioner && !c->prepared) return fail(c, MAS_ERR_INVALID, "mas_prepare first");
	if (c->world > 1) return fail(c, MAS_ERR_UNSUPPORTED, "mas_pcg_solve runs on single-GPU contexts");
	if (c->nnz > 0 && (!csrOffDiagonals || !csrIdx)) return fail(c, MAS_ERR_INVALID, "off-diagonal arrays missing");
	MAS_CUDA(c, cudaSetDevice(c->device));
	const float *dDiag, *dOff; const int *dRanges, *dIdx; const float4* dB;
	if (int rc = stage_in(c, c->pcgDiag, diagonal, (size_t)c->nv * 9, mem, &dDiag)) return rc;
	if (int rc = stage_in(c, c->pcgOff, csrOffDiagonals, (size_t)c->nnz * 9, mem, &dOff)) return rc;
	if (int rc = stage_in(c, c->pcgRanges, csrRanges, (size_t)c->nv + 1, mem, &dRanges)) return rc;
	if (int rc = stage_in(c, c->pcgIdx, csrIdx, (size_t)c->nnz, mem, &dIdx)) return rc;
	if (int rc = stage_in(c, c->pcgB, b, (size_t)c->nv, mem, &dB)) return rc;
	float4* dX = (float4*)x;
	if (mem != MAS_MEM_DEVICE)
	{
		if (int rc = reserve(c, c->pcgX, (size_t)c->nv)) return rc;
		dX = c->pcgX.p;
	}
	if (int rc = pcg_solve(c, dDiag, dOff, dRanges, dIdx, dB, dX, relTol, maxIter, usePreconditioner ? 1 : 0, itersOut, relResOut)) return rc;
	if (mem != MAS_MEM_DEVICE)
	{
		MAS_CUDA(c, cudaMemcpyAsync(x, dX, sizeof(float4) * (size_t)c->nv, cudaMemcpyDeviceToHost, c->stream));
		MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	return MAS_OK;
}

int mas_exchange_buffer(mas_handle_t h, int which, void** device_ptr, size_t* count)
{
	if (!h || !device_ptr || !count) return MAS_ERR_INVALID;
	Context* c = h;
	if (which == 0) { *device_ptr = c->coarseAcc.p; *count = c->coarseAccCount; return MAS_OK; }
	if (which == 1) { *device_ptr = c->coarseR.p; *count = (size_t)c->nCoarseNodes * 4; return MAS_OK; }
	return MAS_ERR_INVALID;
}

int mas_peer_export(mas_handle_t h, void* handle_out_64)
{
	if (!h || !handle_out_64) return MAS_ERR_INVALID;
	Context* c = h;
	if (c->world < 2 || !c->arena.p) return fail(c, MAS_ERR_INVALID, "mas_set_partition(world > 1) and mas_allocate first");
	MAS_CUDA(c, cudaSetDevice(c->device));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t hd;
	MAS_CUDA(c, cudaIpcGetMemHandle(&hd, c->arena.p));
	std::memcpy(handle_out_64, &hd, 64);
	return MAS_OK;
}

int mas_peer_local(mas_handle_t h, void** arena_out)
{
	if (!h || !arena_out) return MAS_ERR_INVALID;
	Context* c = h;
	if (c->world < 2 || !c->arena.p) return fail(c, MAS_ERR_INVALID, "mas_set_partition(world > 1) and mas_allocate first");
	*arena_out = c->arena.p;
	return MAS_OK;
}

int mas_peer_attach(mas_handle_t h, const void* handles, void* const* pointers)
{
	if (!h || (!handles && !pointers)) return MAS_ERR_INVALID;
	Context* c = h;
	if (c->world < 2 || !c->arena.p) return fail(c, MAS_ERR_INVALID, "mas_set_partition(world > 1) and mas_allocate first");
	if (c->world > 16) return fail(c, MAS_ERR_UNSUPPORTED, "peer exchange supports up to 16 ranks");
	MAS_CUDA(c, cudaSetDevice(c->device));
	drop_graph(c);
	close_peers(c);
	for (int q = 0; q < c->world; ++q)
	{
		if (q == c->rank) { c->peerArena[q] = c->arena.p; continue; }
		if (pointers && pointers[q])
		{
			// another context of this process: make sure its device is peer-accessible
			cudaPointerAttributes at;
			MAS_CUDA(c, cudaPointerGetAttributes(&at, pointers[q]));
			if (at.device != c->device)
			{
				cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { check(c, e, "cudaDeviceEnablePeerAccess"); close_peers(c); return MAS_ERR_CUDA; }
				cudaGetLastError();
			}
			c->peerArena[q] = pointers[q];
			continue;
		}
		if (!handles) { close_peers(c); return fail(c, MAS_ERR_INVALID, "no handle or pointer for a peer"); }
		cudaIpcMemHandle_t hd;
		std::memcpy(&hd, (const unsigned char*)handles + 64 * (size_t)q, 64);
		void* p = nullptr;
		if (!check(c, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) { close_peers(c); return MAS_ERR_CUDA; }
		c->peerArena[q] = p;
		c->peerOpened[q] = true;
	}
	if (c->prepared && (size_t)c->nCoarseNodes > c->arenaCap)
	{
		close_peers(c);
		return fail(c, MAS_ERR_UNSUPPORTED, "coarse hierarchy larger than the peer arena");
	}
	c->p2p = true;
	return MAS_OK;
}

int mas_get_int(mas_handle_t h, int key, long long* out)
{
	if (!h || !out) return MAS_ERR_INVALID;
	Context* c = h;
	switch (key)
	{
	case MAS_INT_NUM_VERTS: *out = c->nv; break;
	case MAS_INT_NUM_LEVEL: *out = c->numLevel; break;
	case MAS_INT_TOTAL_CLUSTERS: *out = c->totalClusters; break;
	case MAS_INT_NUM_BLOCKS: *out = c->nBlocks; break;
	case MAS_INT_STENCIL_NUM: *out = c->nStencil; break;
	case MAS_INT_NNZ: *out = c->nnz; break;
	case MAS_INT_APPLY_LAUNCHES: *out = c->applyLaunches; break;
	case MAS_INT_PACKED_FLOATS_PER_BLOCK: *out = kTri; break;
	case MAS_INT_OWNED_BLOCK_BEGIN: *out = c->ownFineBegin; break;
	case MAS_INT_OWNED_BLOCK_END: *out = c->ownFineEnd; break;
	case MAS_INT_PREPARE_LAUNCHES: *out = c->prepareLaunches; break;
	case MAS_INT_PCG_LAUNCHES_PER_ITER: *out = c->pcgLaunchesPerIter; break;
	case MAS_INT_PCG_CONVERGED: *out = c->pcgConverged; break;
	case MAS_INT_ALIGNED_CUTS: *out = c->alignedCuts ? 1 : 0; break;
	case MAS_INT_HOST_PULL_CHOICE: *out = c->pullChoice; break;
	case MAS_INT_HOST_BYTES_IN: *out = c->hostBytesIn; break;
	case MAS_INT_HOST_BYTES_OUT: *out = c->hostBytesOut; break;
	case MAS_INT_PEER_ERROR:
	{
		*out = 0;
		if (c->peerErrHost)
		{
			MAS_CUDA(c, cudaSetDevice(c->device));
			MAS_CUDA(c, cudaStreamSynchronize(c->stream));
			*out = *reinterpret_cast<volatile unsigned*>(c->peerErrHost);
		}
		break;
	}
	default: return fail(c, MAS_ERR_INVALID, "unknown int key");
	}
	return MAS_OK;
}

static int copy_out(Context* c, const void* dev, size_t have, void* host, size_t bytes)
{
	if (bytes > have) return fail(c, MAS_ERR_INVALID, "host buffer larger than the array");
	MAS_CUDA(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
	MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	return MAS_OK;
}

int mas_get_array(mas_handle_t h, int key, int index, void* host_out, size_t bytes)
{
	if (!h || !host_out) return MAS_ERR_INVALID;
	Context* c = h;
	MAS_CUDA(c, cudaSetDevice(c->device));
	const size_t nv = (size_t)c->nv;
	switch (key)
	{
	case MAS_ARR_MORTON: return copy_out(c, c->code.p, 8 * nv, host_out, bytes);
	case MAS_ARR_SORTED_GET_ORIGINAL: return copy_out(c, c->s2o.p, 4 * nv, host_out, bytes);
	case MAS_ARR_ORIGINAL_GET_SORTED: return copy_out(c, c->o2s.p, 4 * nv, host_out, bytes);
	case MAS_ARR_GOING_NEXT: return copy_out(c, c->goingNext.p, 4 * (size_t)c->totalClusters, host_out, bytes);
	case MAS_ARR_LEVEL_SIZE:
	{
		size_t have = sizeof(int) * 2 * (size_t)(c->numLevel + 1);
		if (bytes > have) return fail(c, MAS_ERR_INVALID, "host buffer larger than the array");
		std::memcpy(host_out, c->levelSize, bytes);
		return MAS_OK;
	}
	case MAS_ARR_FINE_CONNECT_MASK: return copy_out(c, c->fineMask.p, 4 * nv, host_out, bytes);
	case MAS_ARR_COARSE_SPACE_TABLE:
		if (index < 0 || index >= c->numLevel) return fail(c, MAS_ERR_INVALID, "level out of range");
		return copy_out(c, c->cst[index].p, 4 * nv, host_out, bytes);
	case MAS_ARR_COARSE_TABLES: return copy_out(c, c->coarseTables.p, 16 * nv, host_out, bytes);
	case MAS_ARR_SORTED_ADJ_STARTS: return copy_out(c, c->adjStart.p, 4 * (nv + 1), host_out, bytes);
	case MAS_ARR_SORTED_ADJ_IDX: return copy_out(c, c->adjIdx.p, 4 * (size_t)c->nnz, host_out, bytes);
	case MAS_ARR_STENCILS: return copy_out(c, c->stencils.p, 80 * (size_t)c->nStencil, host_out, bytes);
	case MAS_ARR_STENCIL_INDEX_MAPPED: return copy_out(c, c->stencilIdx.p, 20 * (size_t)c->nStencil, host_out, bytes);
	case MAS_ARR_DENSE_INVERSE:
		if (index < 0 || index >= c->nBlocks || bytes != sizeof(float) * kDof * kDof) return fail(c, MAS_ERR_INVALID, "bad block / size");
		return unpack_dense_inverse(c, index, (float*)host_out);
	case MAS_ARR_MAPPED_R:
	case MAS_ARR_MAPPED_Z:
	{
		// levels >= 1 only; level-0 entries are not materialised on the device and read as 0
		size_t have = 16 * (size_t)c->totalClusters;
		if (bytes > have) return fail(c, MAS_ERR_INVALID, "host buffer larger than the array");
		std::memset(host_out, 0, bytes);
		size_t skip = 16 * (size_t)c->nVC;
		if (bytes <= skip) return MAS_OK;
		const float4* src = key == MAS_ARR_MAPPED_R ? c->coarseR.p : c->coarseZ.p;
		return copy_out(c, src, 16 * (size_t)c->nCoarseNodes, (char*)host_out + skip, bytes - skip);
	}
	case MAS_ARR_AABB: return copy_out(c, c->aabb.p, 32, host_out, bytes);
	default: return fail(c, MAS_ERR_INVALID, "unknown array key");
	}
}

int mas_morton_encode(mas_handle_t h, const float* xyz, int count, uint64_t* codes_out)
{
	if (!h || !xyz || !codes_out || count <= 0) return MAS_ERR_INVALID;
	MAS_CUDA(h, cudaSetDevice(h->device));
	return morton_encode_points(h, xyz, count, (unsigned long long*)codes_out);
}

int mas_synchronize(mas_handle_t h)
{
	if (!h) return MAS_ERR_INVALID;
	Context* c = h;
	MAS_CUDA(c, cudaSetDevice(c->device));
	MAS_CUDA(c, cudaStreamSynchronize(c->stream));
	return peer_failed(c);
}

int mas_get_timing(mas_handle_t h, int which, float* ms_out)
{
	if (!h || !ms_out) return MAS_ERR_INVALID;
	Context* c = h;
	*ms_out = 0.f;
	if (which == 0) { *ms_out = c->lastPrepareMs; return MAS_OK; }
	if (!c->optTimeKernels) return fail(c, MAS_ERR_INVALID, "set MAS_OPT_TIME_KERNELS first");
	MAS_CUDA(c, cudaSetDevice(c->device));
	if (which == 1)
	{
		MAS_CUDA(c, cudaEventSynchronize(c->evAp1));
		MAS_CUDA(c, cudaEventElapsedTime(ms_out, c->evAp0, c->evAp1));
		return MAS_OK;
	}
	if (which == 2)
	{
		MAS_CUDA(c, cudaEventSynchronize(c->evF1));
		MAS_CUDA(c, cudaEventElapsedTime(ms_out, c->evF0, c->evF1));
		return MAS_OK;
	}
	return fail(c, MAS_ERR_INVALID, "unknown timing key");
}

}  // extern "C"
