/* mas_b200.h — C ABI of the B200-native multilevel additive Schwarz (MAS) preconditioner.
 *
 * Drop-in boundary for the one hot path of
 * V-Sekai/preconditioner-for-cloth-and-deformable-body-simulation:
 * class SE::SeSchwarzPreconditioner (SeSchwarzPreconditioner.h:37-178).
 * The reference defines no FFI; these entry points are exactly what a binding
 * of its three public methods would call.  Plain pointers and sizes only.
 *
 *   reference call (SeSchwarzPreconditioner.h)              replacement
 *   ------------------------------------------------------  ------------------
 *   m_positions/m_edges/m_faces/m_neighbours (h:44-51)
 *   + AllocatePrecoditioner(nv, ne, nf)       (h:56)        mas_allocate()
 *   PreparePreconditioner(diag, offdiag, ranges,
 *        ef, ee, vf, efCounts, eeCounts, vfCounts) (h:59-60) mas_prepare()
 *   Preconditioning(z, residual, dim)          (h:63)       mas_apply()
 *
 * Array layouts are the reference's own:
 *   positions / residual / z : 16-byte xyzw float vectors (SeVec3fSimd, SeVectorSimd.h:45-103)
 *   diag / offdiag           : 36-byte column-major 3x3 floats (SeMatrix3f, SeMatrix.h:681-682)
 *   edges / faces            : 16-byte Int4 rows (h:48-49); only [0],[1] / [0..2] are read
 *   nbrStarts / nbrIdx       : SeCsr<int> m_starts / m_idxs (SeCsr.h:35-173), no self entries;
 *                              offdiag[nbrStarts[i]+k] is A(i, nbrIdx[nbrStarts[i]+k])
 *   ef / ee / vf             : 48-byte EfSet / EeSet / VfSet records (SeCollisionElements.h:33-58),
 *                              indexed the way the reference indexes them (global stencil index, cpp:328/357/383)
 *
 * Every pointer argument is either HOST memory (as in the reference, whose
 * caller owns host arrays) or DEVICE memory on the context's GPU, selected per
 * call by `mem`.  Device pointers avoid the PCIe copies that would otherwise
 * dwarf the apply.  All work is issued on the context's CUDA stream; host-memory
 * calls return after the results are in the caller's buffers, device-memory calls
 * return after enqueueing (mas_prepare synchronises once internally to size
 * buffers from the actual cluster counts).
 *
 * There is no CPU fallback: every entry point fails with MAS_ERR_CUDA if no
 * sm_100 device is usable.
 */
#ifndef MAS_B200_H
#define MAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mas_context* mas_handle_t;

enum
{
	MAS_OK = 0,
	MAS_ERR_INVALID = 1,   /* bad argument / call order */
	MAS_ERR_CUDA = 2,      /* CUDA runtime failure (see mas_last_error) */
	MAS_ERR_UNSUPPORTED = 3 /* e.g. more than 5 levels (Int4 coarse table, cpp:96) */
};

enum
{
	MAS_MEM_HOST = 0,
	MAS_MEM_DEVICE = 1
};

/* mas_set_option keys */
enum
{
	/* 0 (default): reproduce SeSchwarzPreconditioner.cpp:1710 — with 5 levels the top level is
	 * solved but never prolonged.  1: prolong every level. */
	MAS_OPT_PROLONG_ALL_LEVELS = 0,
	/* share of the owned fine domains that the apply graph solves CONCURRENTLY with the coarse-level chain:
	 * -1 (default) sized automatically to the chain's duration, 0 = strictly sequential launches, >0 = per-mille
	 *  */
	MAS_OPT_APPLY_VARIANT = 1,
	/* 1 (default): capture the apply launch sequence in a CUDA graph */
	MAS_OPT_USE_GRAPH = 2,
	/* 1: bracket the dominant apply kernel (level-0 solve) with CUDA events on the launching stream so that
	 * mas_get_timing(h, 2, ..) reports its duration; implies un-captured launches.  Default 0. */
	MAS_OPT_TIME_KERNELS = 3,
	/* sharded contexts, 1 (default): move every shard cut (by at most 1/8 of a shard) to a fine bank where the running
	 * level-1 id is a multiple of 32, so that no level-1 bank straddles two shards and the per-apply exchange carries
	 * level-2 residuals; 0: even split, exchange of level-1 residuals.  Takes effect at the next mas_prepare. */
	MAS_OPT_ALIGN_CUTS = 4,
	/* Q2/Q3 fix mode, default 0 = the reference's literal reading.  1: eeSets and vfSets are read from THEIR
	 * OWN index 0 (the reference indexes all three arrays by the global stencil index, cpp:357/383, so callers must pad
	 * them), and the third VF weight is -(1 - b0 - b1) computed from the two stored barycentrics (the reference reads the
	 * struct's padding float at byte 24 as m_bary[2], cpp:399). */
	MAS_OPT_STENCIL_FIX = 5,
	/* Re-sort policy.  0 (default) = the reference as shipped: the Morton order is built by the first
	 * AllocatePrecoditioner call and never again (m_frameIndex sticks at 1, cpp:44-64, Q1).  N > 0 = the evidently intended
	 * behaviour "re-run SpaceSort every N frames" (the reference hard-codes 17): every N-th mas_allocate call re-reads the
	 * positions and adjacency and rebuilds the ordering; sizes must not change. */
	MAS_OPT_RESORT_PERIOD = 6,
	/* Host-pointer mas_apply (MAS_MEM_HOST) only.  0 (default): the residual crosses PCIe with cudaMemcpyAsync (copy
	 * engine).  1: when the caller's residual buffer is page-locked (cudaHostAlloc / cudaHostRegister, detected with
	 * cudaPointerGetAttributes) a kernel PULLS it through its device mapping with coalesced 16-byte loads instead; pageable
	 * buffers still take the copy engine.  Measured on a virtualised B200 host whose copy engine reads host memory at
	 * 15-20 GB/s while SM loads reach 33 GB/s and D2H runs at 56 GB/s either way (tools/pcie_bound.py); z always returns
	 * through the copy engine.  Results are bit-identical.  2: decide by measurement — the first six host-pointer applies
	 * with a page-locked residual time both stagings on the device (three each), the faster one is kept
	 * (mas_get_int(MAS_INT_HOST_PULL_CHOICE): -1 undecided, 0 copy engine, 1 kernel pull). */
	MAS_OPT_HOST_PULL = 7,
	/* Batched 96x96 inversion kernel.  0 (default): tcgen05 tensor cores — block Gauss-Jordan by 16-column panels, every
	 * panel update one rank-16 GEMM with 3xTF32 operands (hi/lo split, FP32 accumulation in tensor memory), which holds the
	 * parity tolerance where plain TF32 misses it by three orders of magnitude.  1: the FP32 CUDA-core
	 * kernel (the reference's LDL^T elimination regrouped by 16x16 tiles).  Takes effect at the next mas_prepare. */
	MAS_OPT_INVERT_VARIANT = 8,
	/* Host-pointer mas_apply only, default 0.  1: a PAGEABLE residual / z buffer (std::vector, malloc) is page-locked where it
	 * lies with cudaHostRegister the first time it is seen, so that every later copy runs at pinned-memory speed; up to
	 * four ranges per context, unlocked by mas_destroy or by setting the option back to 0.  The caller must keep such a
	 * buffer allocated until then (the reference's callers reuse r and z for the whole solve). */
	MAS_OPT_REGISTER_HOST = 9,
	/* Incremental setup, default 1: a PreparePreconditioner without collision stencils that follows another
	 * one without stencils keeps the clustering (levels, goingNext, coarse tables, shard cuts): it depends on the sorted
	 * adjacency and the stencils only, so the rebuilt one would be identical bit for bit (verified on hardware).  Assembly
	 * and inversion always run.  A prepare with stencils, a re-sort or MAS_OPT_ALIGN_CUTS rebuilds.  0: always rebuild. */
	MAS_OPT_CACHE_HIERARCHY = 10,
	/* Sharded apply over peer memory, default 0.  The kernel that publishes a rank's coarse residuals stores its arrival flags
	 * with st.relaxed.sys right after the kernel that wrote the payload has finished (the payload then sits in this GPU's L2,
	 * the point of coherence for reads arriving over NVLink).  1: a system-scope fence precedes the flag stores, which is what
	 * the PTX memory model guarantees between GPUs (kernel boundaries order at device scope); it costs a MEMBAR.SYS on the
	 * latency-bound chain.  Bit-identical results either way on current NVLink hardware. */
	MAS_OPT_STRICT_PUBLISH = 11,
	/* mas_pcg_solve only, default 1: for the duration of a solve the vectors of the iteration (r, z, p, Ap: 64 MB at 1M
	 * vertices) are kept in L2 through a persisting access-policy window on every kernel of the iteration graph, while the
	 * matrix and the packed inverses stream past them; the L2 carve-out (cudaLimitPersistingL2CacheSize, a per-device
	 * setting; the value found is restored) is released when the solve returns.  0: no window, no carve-out.  Results are bit-identical either way. */
	MAS_OPT_PCG_PERSIST_L2 = 12
};

/* mas_get_int keys */
enum
{
	MAS_INT_NUM_VERTS = 0,
	MAS_INT_NUM_LEVEL = 1,       /* m_numLevel (h:74) */
	MAS_INT_TOTAL_CLUSTERS = 2,  /* m_totalNumberClusters (h:76) */
	MAS_INT_NUM_BLOCKS = 3,      /* active 32-node domains over all levels */
	MAS_INT_STENCIL_NUM = 4,     /* m_stencilNum (h:112) */
	MAS_INT_NNZ = 5,
	MAS_INT_APPLY_LAUNCHES = 6,  /* kernels launched per mas_apply */
	MAS_INT_PACKED_FLOATS_PER_BLOCK = 7,
	MAS_INT_OWNED_BLOCK_BEGIN = 8,
	MAS_INT_OWNED_BLOCK_END = 9,
	MAS_INT_PREPARE_LAUNCHES = 10, /* kernels launched by the last mas_prepare */
	MAS_INT_PCG_LAUNCHES_PER_ITER = 11, /* kernels per iteration of the last mas_pcg_solve (its own 4 + the apply's) */
	MAS_INT_PCG_CONVERGED = 12,    /* 1 if the last mas_pcg_solve met its tolerance */
	MAS_INT_PEER_ERROR = 13,       /* 1 if a peer-memory wait ever timed out (a rank stopped publishing) */
	MAS_INT_ALIGNED_CUTS = 14,     /* sharded contexts: 1 if no level-1 bank straddles a shard cut (the apply exchanges level-2 residuals) */
	MAS_INT_HOST_PULL_CHOICE = 15, /* MAS_OPT_HOST_PULL = 2: -1 still sampling, 0 copy engine kept, 1 kernel pull kept */
	MAS_INT_HOST_BYTES_IN = 16,    /* bytes the last host-pointer mas_apply moved host -> device ... */
	MAS_INT_HOST_BYTES_OUT = 17    /* ... and device -> host (a shard with page-locked buffers moves its own vertices only) */
};

/* mas_get_array keys: copies an internal device array to a HOST buffer (parity tests) */
enum
{
	MAS_ARR_MORTON = 0,              /* uint64[nv]   m_mortonCode (h:105), original order */
	MAS_ARR_SORTED_GET_ORIGINAL = 1, /* int32[nv]    m_MapperSortedGetOriginal (h:103) */
	MAS_ARR_ORIGINAL_GET_SORTED = 2, /* int32[nv]    m_mapperOriginalGetSorted (h:104) */
	MAS_ARR_GOING_NEXT = 3,          /* int32[totalClusters] m_goingNext (h:97) */
	MAS_ARR_LEVEL_SIZE = 4,          /* int32[(numLevel+1)*2] m_levelSize (h:99) */
	MAS_ARR_FINE_CONNECT_MASK = 5,   /* uint32[nv]   m_fineConnectMask after PreparePrefixSumL0 (h:90) */
	MAS_ARR_COARSE_SPACE_TABLE = 6,  /* int32[nv] for level `index`: m_CoarseSpaceTables[index] (h:88) */
	MAS_ARR_COARSE_TABLES = 7,       /* int32[nv*4]  m_coarseTables (h:96); entries >= numLevel-1 are 0 */
	MAS_ARR_SORTED_ADJ_STARTS = 8,   /* int32[nv+1]  adjacency in sorted space (m_mappedNeighbors, h:85, as CSR) */
	MAS_ARR_SORTED_ADJ_IDX = 9,      /* int32[nnz] */
	MAS_ARR_STENCILS = 10,           /* 80-byte Stencil records [stencilNum] (SeCollisionElements.h:60-69) */
	MAS_ARR_STENCIL_INDEX_MAPPED = 11, /* int32[stencilNum*5] m_stencilIndexMapped (h:115) */
	MAS_ARR_DENSE_INVERSE = 12,      /* float[96*96] dense symmetric inverse of 32-node block `index` */
	MAS_ARR_MAPPED_R = 13,           /* float[totalClusters*4] m_mappedR (h:101) after the last apply (levels >= 1 only; level 0 is not materialised and reads 0) */
	MAS_ARR_MAPPED_Z = 14,           /* float[totalClusters*4] m_mappedZ (h:102), levels >= 1 only */
	MAS_ARR_AABB = 15                /* float[8] lower xyzw, upper xyzw (h:79) */
};

int mas_create(mas_handle_t* out, int device);
int mas_destroy(mas_handle_t h);
const char* mas_last_error(mas_handle_t h);

/* cudaStream_t as void*; NULL = the legacy default stream */
int mas_set_stream(mas_handle_t h, void* cuda_stream);
int mas_set_option(mas_handle_t h, int key, int value);

/* Shard the 32-node fine domains across `world` GPUs in Morton-contiguous ranges
 * (SURVEY §8e).  Must precede mas_allocate.  Default: rank 0 of 1.  At most 16 ranks (MAS_ERR_UNSUPPORTED beyond). */
int mas_set_partition(mas_handle_t h, int rank, int world);

/* replaces m_positions/m_edges/m_faces/m_neighbours + AllocatePrecoditioner (h:44-56, cpp:38-65).
 * One mesh per handle, as in the reference (its second call skips the allocation and the sort, cpp:44-64): a later call with
 * the same sizes is a no-op (or a re-sort, MAS_OPT_RESORT_PERIOD); a call with OTHER sizes returns MAS_ERR_INVALID instead of
 * carrying on with the first mesh's buffers. */
int mas_allocate(mas_handle_t h, int numVerts, int numEdges, int numFaces,
	const float* positions, const int* edges, const int* faces,
	const int* nbrStarts, const int* nbrIdx, int mem);

/* replaces PreparePreconditioner (h:59-60, cpp:67-98).  efTotal/eeTotal/vfTotal are the values the
 * reference reads from efCounts[numEdges], eeCounts[numEdges], vfCounts[numVerts] (cpp:306-308). */
int mas_prepare(mas_handle_t h, const float* diagonal, const float* csrOffDiagonals, const int* csrRanges,
	const void* efSets, const void* eeSets, const void* vfSets,
	unsigned efTotal, unsigned eeTotal, unsigned vfTotal, int mem);

/* replaces Preconditioning (h:63, cpp:100-110); `dim` is unused there and dropped here.  z == residual (in place) works with
 * host pointers as it does in the reference; with device pointers overlapping z and residual return MAS_ERR_INVALID. */
int mas_apply(mas_handle_t h, float* z, const float* residual, int mem);

/* Multi-GPU (world > 1) phase split.  Between *_begin and *_end the caller sums the exchange
 * buffer across ranks (one all-reduce over NCCL):
 *   prepare: coarse-level Galerkin accumulators (FP64);  apply: coarse-level residuals (FP32). */
int mas_prepare_begin(mas_handle_t h, const float* diagonal, const float* csrOffDiagonals, const int* csrRanges,
	const void* efSets, const void* eeSets, const void* vfSets,
	unsigned efTotal, unsigned eeTotal, unsigned vfTotal, int mem);
int mas_prepare_end(mas_handle_t h);
int mas_apply_begin(mas_handle_t h, const float* residual, int mem);
int mas_apply_end(mas_handle_t h, float* z, int mem);
/* which: 0 = prepare exchange (double), 1 = apply exchange (float). Returns a DEVICE pointer + element count. */
int mas_exchange_buffer(mas_handle_t h, int which, void** device_ptr, size_t* count);

/* Multi-GPU apply exchange over peer memory (NVLink / NVSwitch), replacing the all-reduce of mas_apply_begin/_end:
 * every rank's level-0 restriction kernel stores its level-1 residuals straight into an arena of every other rank and
 * raises a flag there; the coarse levels wait on those flags on the device.  With the peers attached, mas_apply() on
 * a sharded context is again ONE call (one CUDA graph, no host synchronisation, no NCCL call per apply).
 *   mas_peer_export : 64-byte cudaIpcMemHandle_t of this rank's arena, to be sent to the other PROCESSES
 *   mas_peer_local  : the arena's device pointer, for other contexts of THIS process
 *   mas_peer_attach : handles = world x 64 bytes in rank order (or NULL), pointers = world device pointers (or NULL);
 *                     for each peer the pointer is used if non-NULL, else the handle is opened.
 * Call after mas_allocate on every rank (the arena is sized there).  Setup still uses mas_prepare_begin/_end. */
int mas_peer_export(mas_handle_t h, void* handle_out_64);
int mas_peer_local(mas_handle_t h, void** arena_out);
int mas_peer_attach(mas_handle_t h, const void* handles, void* const* pointers);

/* Caller-side harness (the reference ships no solver): preconditioned conjugate gradients for A x = b with
 * everything resident on the GPU.  A is given exactly as PreparePreconditioner receives it (original vertex order):
 * diagonal[nv] and csrOffDiagonals[nnz] 36-byte column-major blocks, csrRanges[nv+1] / csrIdx[nnz] the adjacency CSR
 * (h:51 m_neighbours); the Hessians of the collision stencils of the last mas_prepare, stiff (w (x) w) (x) (d d^T) per stencil
 * (cpp:1201-1227) — which the preconditioner was built for — are part of A and applied matrix-free.  b, x: 16-byte xyzw vectors, x0 = 0.  Stops when ||r||_2 / ||b||_2 < relTol or after maxIter
 * iterations; dot products in FP64.  usePreconditioner = 0 runs plain CG (z = r).  Requires mas_prepare when
 * usePreconditioner != 0.  Single-GPU contexts only. */
int mas_pcg_solve(mas_handle_t h, const float* diagonal, const float* csrOffDiagonals, const int* csrRanges, const int* csrIdx,
	const float* b, float* x, float relTol, int maxIter, int usePreconditioner, int mem, int* itersOut, float* relResOut);

/* Waits for everything enqueued on the context's stream.  On a sharded context this is also where a failed peer exchange
 * surfaces: if a device-side wait for another rank's coarse residuals ever timed out (~2 s; that rank died or never
 * launched its apply) this call — and every later mas_apply — returns MAS_ERR_CUDA; the z of that apply is invalid. */
int mas_synchronize(mas_handle_t h);

/* introspection */
int mas_get_int(mas_handle_t h, int key, long long* out);
int mas_get_array(mas_handle_t h, int key, int index, void* host_out, size_t bytes);

/* stand-alone helper with the exact bit behaviour of SeMorton64::Encode (SeMorton.h:75-86), evaluated on the GPU */
int mas_morton_encode(mas_handle_t h, const float* xyz, int count, uint64_t* codes_out);

/* device-side CUDA-event timing in milliseconds: which = 0 last mas_prepare (whole call), 1 last mas_apply
 * (whole call, only with MAS_OPT_TIME_KERNELS), 2 level-0 solve kernel of the last mas_apply (only with
 * MAS_OPT_TIME_KERNELS).  Synchronises the stream. */
int mas_get_timing(mas_handle_t h, int which, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* MAS_B200_H */
