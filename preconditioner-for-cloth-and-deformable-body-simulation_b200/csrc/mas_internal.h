// Internal definitions shared by the CUDA translation units of libmas_b200.so.
// Not part of the public boundary (include/mas_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/mas_b200.h"

namespace mas {

constexpr int kBank = 32;          // nodes per domain (cpp:139)
constexpr int kDof = 96;           // 3 * kBank
constexpr int kTri = 4656;         // 96*97/2 unique floats of a symmetric 96x96 block
constexpr int kMaxLevel = 5;       // Int4 coarse table (h:96) => at most 5 levels
constexpr int kMaxCollisionPerVert = 32;  // cpp:187

// ---------------------------------------------------------------------------
// Packed inverse layout ("lane-slot" layout), 4656 floats = 18,624 bytes per
// 32-node domain, no padding.  A domain matrix is a 32x32 grid of 3x3 blocks;
// one warp applies it with lane i owning node i.  Lane i stores
//   for d = 1..15 : the full 3x3 block A(i, (i+d)%32)            (9 floats each)
//   then          : the 6 unique floats of its diagonal block A(i,i)
// = 141 floats, laid out as 35 float4 "slots" [q][lane] (each slot is one fully
// coalesced 512-byte warp load) plus one trailing float [lane].  The 16 blocks
// A(i, i+16), i < 16, follow as 2 float4 slots + 1 float over 16 lanes.
//   floats [   0,4480) : slot q (0..34), lane l  -> 4*(32*q + l) + e,  e = n%4, q = n/4, n < 140
//   floats [4480,4512) : n = 140 (a22 of the diagonal block), lane l
//   floats [4512,4640) : half-diagonal, slot q (0..1), lane l<16 -> 4512 + 4*(16*q + l) + e
//   floats [4640,4656) : half-diagonal float 8, lane l<16
// Per-lane index n: n = 9*(d-1) + 3*a + b for block d entry (a,b); n = 135 + k6
// for the diagonal block with k6 = a*(a+1)/2 + b (a >= b).
// ---------------------------------------------------------------------------
constexpr int kLaneFloats = 141;
constexpr int kFullSlots = 35;
constexpr int kTailBase = 4480;
constexpr int kHalfBase = 4512;
constexpr int kHalfTail = 4640;

__host__ __device__ inline int packed_lane_pos(int lane, int n)
{
	return n < 140 ? 4 * (32 * (n >> 2) + lane) + (n & 3) : kTailBase + lane;
}
__host__ __device__ inline int packed_half_pos(int lane, int n)
{
	return n < 8 ? kHalfBase + 4 * (16 * (n >> 2) + lane) + (n & 3) : kHalfTail + lane;
}
// position of symmetric entry (r,c) of the 96x96 block, any order of r,c
__host__ __device__ inline int packed_pos(int r, int c)
{
	int i = r / 3, a = r - 3 * i, j = c / 3, b = c - 3 * j;
	if (i == j)
	{
		if (a < b) { int t = a; a = b; b = t; }
		return packed_lane_pos(i, 135 + a * (a + 1) / 2 + b);
	}
	int d = (j - i) & 31;
	if (d == 16)
	{
		if (i > j) { int t = i; i = j; j = t; t = a; a = b; b = t; }
		return packed_half_pos(i, 3 * a + b);
	}
	if (d > 16) { int t = i; i = j; j = t; t = a; a = b; b = t; d = 32 - d; }
	return packed_lane_pos(i, 9 * (d - 1) + 3 * a + b);
}

struct Stencil  // SeCollisionElements.h:60-69 byte layout (80 bytes, 16-aligned)
{
	int n, nFirst;
	int index[5];
	float weight[5];
	float stiff;
	float pad_[3];
	float dir[4];
};
static_assert(sizeof(Stencil) == 80, "Stencil must match the reference layout");

template <typename T>
struct DevBuf
{
	T* p = nullptr;
	size_t cap = 0;  // elements
};

struct Context
{
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	int optProlongAll = 0;
	int optApplyVariant = -1;    // fine banks solved concurrently with the coarse chain: -1 auto, 0 none, >0 per-mille of the owned banks
	int optUseGraph = 1;
	int optTimeKernels = 0;
	int optAlignCuts = 1;
	int optStencilFix = 0;       // read eeSets / vfSets from their own index 0 and form the third VF weight from b0 + b1 (Q2/Q3 fixed)
	int optResortPeriod = 0;     // > 0: every that many mas_allocate calls the Morton order is rebuilt (0: once per object, Q1)
	int optInvertVariant = 0;    // 0: tcgen05 tensor-core inversion (3xTF32 block Gauss-Jordan); 1: FP32 CUDA-core blocked LDL^T
	int optHostPull = 0;         // host-pointer apply: 1 pull a page-locked residual with a kernel instead of the copy engine, 2 pick the faster
	int optPcgPersistL2 = 1;     // mas_pcg_solve: persisting L2 window over the vectors of the iteration
	int optStrictPublish = 0;    // 1: system-scope fence before the peer flag stores of the sharded apply
	int optCacheHierarchy = 1;   // 1: a collision-free prepare reuses the clustering of the previous collision-free prepare
	bool hierarchyCached = false;   // the hierarchy in this context was built without stencils for the current ordering / options
	int optRegisterHost = 0;     // host-pointer apply: page-lock the caller's pageable r / z in place (cudaHostRegister) on first sight
	struct HostRange { const void* p = nullptr; size_t bytes = 0; };
	HostRange registered[4];     // ranges this context page-locked (and must unlock)
	long long hostBytesIn = 0, hostBytesOut = 0;   // PCIe bytes of the last host-pointer apply (host -> device, device -> host)
	int pullCalls = 0;           // auto mode: host-pointer applies sampled so far (3 per staging, the first of each is warm-up)
	int pullChoice = -1;         // auto mode: -1 undecided, 0 copy engine, 1 kernel pull
	float pullBestMs[2] = { 1e30f, 1e30f };
	int allocateCalls = 0;
	int rank = 0, world = 1;
	int smCount = 148;

	// ---- allocate-time state
	bool allocated = false;
	int nv = 0, ne = 0, nf = 0, nnz = 0, nVC = 0, numLevel = 0;
	DevBuf<float4> positions;       // staging when the caller passes host memory
	DevBuf<int4> edges, faces;
	DevBuf<int> inStarts, inIdx;    // caller adjacency (original space)
	DevBuf<float> aabb;             // 8 floats
	DevBuf<unsigned long long> code, codeSorted;
	DevBuf<int> s2o, o2s, iota;
	DevBuf<int> adjStart, adjIdx;   // adjacency in sorted space (CSR)
	DevBuf<unsigned char> cubTemp;

	// ---- prepare-time state
	bool prepared = false;
	int nStencil = 0;
	DevBuf<Stencil> stencils;
	DevBuf<int> stencilIdx;          // [nStencil][5] sorted-space ids
	DevBuf<int> stencilFlag, stencilSlot;
	DevBuf<unsigned> fineMask;       // [nVC]
	DevBuf<int> cst[kMaxLevel];      // CoarseSpaceTables[level][nv]
	DevBuf<int> goingNext;
	DevBuf<unsigned> nextMask;
	DevBuf<int> nextId;
	DevBuf<int> bankCount, bankPrefix;
	DevBuf<int> scanTotal;           // 1 int
	DevBuf<int4> coarseTables;
	int levelSize[kMaxLevel + 2][2] = {};
	int totalClusters = 0;
	int nBlocks = 0;       // totalClusters / 32
	int nFineBlocks = 0;   // nVC / 32
	int nCoarseNodes = 0;  // totalClusters - nVC

	DevBuf<float> diagIn, offdiagIn;     // staging for host inputs
	DevBuf<int> rangesIn;
	DevBuf<unsigned char> efIn, eeIn, vfIn;
	DevBuf<float> extraFine;             // [nv][9] collision self terms (m_additionalHessian32 for vertices)
	DevBuf<int> cooCount, cooStart, cooFill;  // per fine bank: level-0 collision pair entries
	DevBuf<float> cooVal;                // [entries][10]: packed (row,col) + 9 floats
	DevBuf<double> coarseAcc;            // exchange buffer: [nCoarseBlocks][96*96] dense + [nCoarseNodes][9] carry
	size_t coarseAccCount = 0;
	DevBuf<float> packedInv;             // [nBlocks][kTri]
	DevBuf<unsigned short> posTab;       // packed positions of the CUDA-core inversion kernel's register-tile outputs
	DevBuf<unsigned short> posTab96;     // tensor-core inversion kernel: packed position of (r, c), r >= c
	DevBuf<int> inputErr;                // raised by the input checks (adjacency / edge / face / stencil ids beyond the mesh)
	DevBuf<int> invertErr;               // set by the tensor-core kernel if an MMA completion wait timed out

	// ---- apply-time state
	DevBuf<float4> coarseR, coarseZ, coarseZsum;  // indexed by node - nVC
	DevBuf<float4> rIn, zOut;                      // staging for host r / z
	// instantiated apply graphs, one per (residual, z) pointer pair the caller uses (a solver may precondition more than one
	// vector pair per iteration); least recently used slot is recycled
	struct ApplyGraphSlot
	{
		cudaGraphExec_t exec = nullptr;
		const float* r = nullptr;
		float* z = nullptr;
		unsigned long long lastUse = 0;
	};
	static constexpr int kApplyGraphSlots = 4;
	ApplyGraphSlot applyGraphs[kApplyGraphSlots];
	unsigned long long applyGraphClock = 0;
	const float* phaseR = nullptr;   // residual of the running mas_apply_begin / _end pair
	int applyLaunches = 0;
	int prepareLaunches = 0;

	// ---- PCG harness workspace (mas_pcg.cu)
	DevBuf<float4> pcgR, pcgB, pcgX;   // pcgR: r, z, p, Ap in one allocation
	DevBuf<double> pcgPartials;
	DevBuf<unsigned char> pcgState;
	DevBuf<float> pcgDiag, pcgOff;
	DevBuf<int> pcgRanges, pcgIdx;
	DevBuf<int> pcgSliceSlots, pcgSliceStart, pcgEllIdx;   // sliced-ELL copy of A built by every mas_pcg_solve
	DevBuf<float> pcgEllVal;
	int pcgLaunchesPerIter = 0;
	int pcgConverged = 0;

	// peer-memory exchange (world > 1): arena = [send0][send1][flags kMaxWorld][epoch][-][error]
	DevBuf<unsigned char> arena;
	size_t arenaCap = 0;              // float4 elements per receive buffer
	void* peerArena[16] = {};         // arena base of every rank, peer-mapped; [rank] is the local one
	bool peerOpened[16] = {};         // opened through cudaIpcOpenMemHandle (must be closed)
	bool p2p = false;
	bool phaseSplit = false;          // inside mas_apply_begin / mas_apply_end: the caller does the exchange, peers are ignored
	unsigned* peerErrHost = nullptr;  // page-locked, device-mapped word (allocated by mas_create): 1 = a peer wait timed out (sticky)
	unsigned* peerErrDev = nullptr;   // its device address

	// partition (fine banks owned by this rank)
	int ownFineBegin = 0, ownFineEnd = 0;
	int l1Slice[17] = {};              // level-1 nodes (level-local ids) produced by rank q: [l1Slice[q], l1Slice[q+1])
	// Coarse blocks this rank inverts and solves: the level-1 blocks [l1BlockBegin, l1BlockEnd) that hold its own level-1
	// nodes (a block straddling two shards is done by both) plus every block of levels >= 2, [nL1Blocks, nCoarseBlocks).
	int l1BlockBegin = 0, l1BlockEnd = 0, nL1Blocks = 0;
	// Shard cuts are moved (within +-1/8 of a shard) to fine banks where the running level-1 id is a multiple of 32, so that
	// no level-1 bank straddles two shards.  When every cut could be aligned, a rank restricts its own level-1 banks to
	// level 2 by itself and the per-apply exchange shrinks to the level-2 residuals (nv/1024 nodes instead of nv/32);
	// l2Slice[q] = first level-2 node (level-local id) produced by rank q's level-1 banks.
	bool alignedCuts = false;
	int l2Slice[17] = {};
	DevBuf<int> cutInfo;               // [3 * 16]: chosen cut bank, aligned flag, level-1 id at the cut

	cudaEvent_t evA = nullptr, evB = nullptr;      // prepare
	cudaEvent_t evAp0 = nullptr, evAp1 = nullptr;  // whole apply (timed mode)
	cudaEvent_t evF0 = nullptr, evF1 = nullptr;    // level-0 solve kernel (timed mode)
	cudaEvent_t evS0 = nullptr, evS1 = nullptr;    // residual staging of the host-pointer apply (MAS_OPT_HOST_PULL = 2)
	cudaStream_t sideA = nullptr, sideB = nullptr; // branches of the apply graph
	cudaEvent_t evFork = nullptr, evHead = nullptr, evCoarse = nullptr, evTail = nullptr;
	float lastApplyMs = 0.f, lastPrepareMs = 0.f;
};

// ---- helpers ---------------------------------------------------------------
bool check(Context* c, cudaError_t e, const char* what);
#define MAS_CUDA(c, call)                         \
	do {                                          \
		if (!::mas::check((c), (call), #call)) return MAS_ERR_CUDA; \
	} while (0)

template <typename T>
int reserve(Context* c, DevBuf<T>& b, size_t n)
{
	if (n <= b.cap && b.p) return MAS_OK;
	if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
	size_t want = n ? n : 1;
	if (!check(c, cudaMalloc((void**)&b.p, want * sizeof(T)), "cudaMalloc")) return MAS_ERR_CUDA;
	b.cap = want;
	return MAS_OK;
}
template <typename T>
void release(DevBuf<T>& b)
{
	if (b.p) cudaFree(b.p);
	b.p = nullptr;
	b.cap = 0;
}

// scratch allocation local to one call: freed on every return path
template <typename T>
struct TempBuf : DevBuf<T>
{
	TempBuf() = default;
	TempBuf(const TempBuf&) = delete;
	TempBuf& operator=(const TempBuf&) = delete;
	~TempBuf() { release(*this); }
};

inline int pad32(int x) { return (x + 31) / 32 * 32; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- stages (each in its own translation unit) -----------------------------
int order_vertices(Context* c, const float4* positions, const int* inStarts, const int* inIdx);  // mas_order.cu
int morton_encode_points(Context* c, const float* xyz, int count, unsigned long long* out);       // mas_order.cu
int build_stencils(Context* c, const void* ef, const void* ee, const void* vf, unsigned efN, unsigned eeN, unsigned vfN);  // mas_cluster.cu
int launch_exclusive_scan(Context* c, const int* in, int count, int* out, int* totalOut);                    // mas_cluster.cu
int build_hierarchy(Context* c);                                                                  // mas_cluster.cu
int assemble_and_invert_begin(Context* c, const float* diag, const float* offdiag, const int* ranges);  // mas_assemble.cu
int assemble_and_invert_end(Context* c);                                                          // mas_assemble.cu
int unpack_dense_inverse(Context* c, int block, float* hostOut);                                  // mas_assemble.cu
int apply_begin(Context* c, const float4* r);                                                     // mas_apply.cu
int apply_end(Context* c, const float4* r, float4* z);                                            // mas_apply.cu
int apply_forked(Context* c, const float4* r, float4* z, cudaStream_t st);                        // mas_apply.cu
int prioritize_apply_graph(Context* c, cudaGraph_t graph);                                        // mas_apply.cu
inline bool use_peers(const Context* c) { return c->p2p && !c->phaseSplit; }
int pcg_solve(Context* c, const float* diag, const float* off, const int* ranges, const int* idx, const float4* b, float4* x,
	float relTol, int maxIter, int usePrecond, int* itersOut, float* relResOut);                  // mas_pcg.cu

}  // namespace mas

struct mas_context : public mas::Context {};
