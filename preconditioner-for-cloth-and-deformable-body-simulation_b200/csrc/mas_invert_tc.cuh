// Batched 96x96 inversion on the Blackwell tensor cores (tcgen05.mma kind::tf32, accumulator in tensor memory): the default
// setup path.  Included once by mas_assemble.cu inside namespace mas { namespace { ... } }, after PhaseClock.
// Replaces LDLtInverse512 (SeSchwarzPreconditioner.cpp:1347-1546): same result within the parity bar, different algorithm.
//
// Algorithm: block Gauss-Jordan ("sweep") inversion of the symmetric positive definite system by panels of 16 columns.
// With T = A at the start, panel K (columns 16K .. 16K+15), C = T[:, K] (96 x 16) and P = T[K, K]^-1 (16 x 16):
//     T[i, j] -= (C P)[i, :] . C[j, :]            for i, j outside block K
//     T[i, K]  = (C P)[i, :],   T[K, j] = (C P)[j, :]^T,   T[K, K] = -P
// and after the six panels T = -A^-1.  T stays symmetric throughout, the swept blocks hold (minus) inverses of principal
// submatrices and the others Schur complements, so no entry outgrows ||A|| or ||A^-1||; there is no separate E^T D^-1 E
// product as in the reference's LDL^T route.  Everything a panel does to T is ONE rank-16 GEMM over the full 96 x 96 square,
//     T += Aop . Bop^T,    Aop[i] = -(C P)[i],  Bop[j] = C[j]            (i, j outside K)
//                          Aop[x] = P[x, :],    Bop[x] = -I[x, :]        (x in K; row and column block K of T zeroed first)
// which is what the tensor core is for: M = 128 (96 used) x N = 96 x K = 16 per panel, accumulated in place in TMEM.
//
// The product C P is a GEMM as well (96 x 16 x 16) and runs on the tensor core first: D[:, K] = C P^T lands in column block K
// of T itself — exactly where (C P) has to end up — with the B operand buffer (the rows of C) as its A operand; the rows then
// read their (C P)[i] back from tensor memory to form Aop, and the update GEMM leaves column block K alone (Bop[x] = 0 for
// x in K; the 16 rows of block K store -P into it).  That took 340 of a row warp's 540 instructions per panel off the CUDA
// cores (ncu: the kernel was at 43 % issue utilisation, FMA pipe 17 %, tensor pipe 13 %).
//
// Precision: kind::tf32 keeps 10 mantissa bits of each operand, which misses the parity bar by three orders of magnitude;
// every operand is split into hi + lo TF32 halves and the three products lo*hi + hi*lo + hi*hi are accumulated in FP32
// (3xTF32: six tcgen05.mma per panel), which is indistinguishable from FP32 arithmetic on the oracle's blocks at
// k/m = 10 .. 1e5 (tools/sweep_inversion_study.py replays this file's algorithm in numpy).  The 16x16 pivot
// inverses, the products C P and all bookkeeping stay FP32 on the CUDA cores.
//
// Mapping: one CTA of 128 threads per system, persistent over systems, FIVE CTAs per SM: a CTA's serial chain — TMEM load,
// pivot inverse, C P, operand store, MMA — is hidden behind the other four.  Tensor memory is allocated in powers of two, so
// each CTA takes 64 + 32 of the SM's 512 columns (5 x 96 = 480) and every panel update is issued as two MMAs, N = 64 and
// N = 32 (same tensor-pipe time as one N = 96); with one 128-column allocation only four CTAs fit (measured: 1.98 -> ms).
// Thread r < 96 owns row r of T: TMEM lane r, read and written with the 32-lane x 32-bit shape, so the column panel C[r, :]
// is one tcgen05.ld of 16 columns and (C P)[r, :] is thread-local.  Warp 3 has no rows: it inverts the pivot block in
// registers (lanes = rows, shuffles broadcast the pivot row) while warps 0-2 zero row / column block K, and its lane 0 issues
// the MMAs.  Operands are written to shared memory in the K-major no-swizzle core-matrix layout (mas_tcgen05.cuh).
constexpr int kTcThreads = 128;
constexpr int kTcColsA = 64, kTcColsB = 32;   // TMEM columns per CTA: matrix columns 0..63 and 64..95 (two allocations)
constexpr int kTcPs = 20;               // row stride of the 16x16 pivot scratch (16-byte aligned rows)

struct TcOperands                       // 28 KB, K-major core-matrix layout, rows 96..127 of A stay zero
{
	float aHi[128 * 16], aLo[128 * 16], bHi[96 * 16], bLo[96 * 16];
};
struct TcSmem
{
	union
	{
		alignas(128) float A[kDof * kLdP];      // the assembled system, row stride 97 (assembly; read once into TMEM)
		TcOperands op;                          // panel operands (elimination)
		float packed[kTri];                     // packed inverse (epilogue)
	};
	alignas(16) float piv[16 * kTcPs];          // pivot block T[K, K]
	alignas(16) float P[16 * kTcPs];            // its inverse (FP32, row stride kTcPs)
	union
	{
		struct                                  // assembly only, as in InvSmem
		{
			float ownDiag[kBank][9];
			int parent[kBank];
			float fold[kBank][9];
		};
		struct { alignas(128) float pHi[16 * 16], pLo[16 * 16]; };   // elimination: P as a tensor-core operand (hi / lo halves)
	};
	alignas(8) uint64_t bar;                    // mbarrier: completion of a panel's MMAs
	uint32_t tmemBase[2];                       // columns 0..63 and 64..95
	int nextWork;                               // next system of this CTA (dynamic distribution)
};
// tensor-memory addresses of a CTA: matrix column c lives at colBase(c) (+ lane << 16)
struct TcAddr
{
	uint32_t a, b;
	__device__ __forceinline__ uint32_t col(int c) const { return c < kTcColsA ? a + (uint32_t)c : b + (uint32_t)(c - kTcColsA); }
};

// In-place un-pivoted block Gauss-Jordan inverse of the 16x16 pivot block, one warp, 2x2 pivots.  Lanes l and l + 16 hold
// the two halves (8 columns each) of row l & 15.  Step q takes the pivot pair P = {2q, 2q+1} with B = M[P, P]:
//   rows of P:   M[P, j] <- B^-1 M[P, j] (j outside P),  M[P, P] <- B^-1
//   other rows:  g = M[i, P] B^-1;  M[i, j] -= g M[P, j] (j outside P),  M[i, P] <- -g
// i.e. two steps of the scalar elimination at once.  The eight steps are ONE dependent chain (shuffle -> determinant ->
// reciprocal -> two FMAs deep), the serial part of every panel, so its length is what counts: the 2x2 pivots halve the number
// of shuffle and reciprocal latencies against sixteen scalar steps (measured: 2.4 k -> cycles per panel), there is no
// divergent branch (a pivot row is the same update with base 0 and g = -B^-1 row) and the reciprocal is MUFU + one Newton step.
__device__ __forceinline__ float rcp_newton(const float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return __fmaf_rn(r, __fmaf_rn(-x, r, 1.0f), r);
}
__device__ __forceinline__ void invert16_warp(const float* __restrict__ piv, float* __restrict__ P, float* __restrict__ pHi,
	float* __restrict__ pLo, const int lane)
{
	constexpr unsigned kAll = 0xffffffffu;
	const int row = lane & 15, half = lane >> 4, c0 = 8 * half;
	float m[8];
	{
		const float4 a = *reinterpret_cast<const float4*>(piv + row * kTcPs + c0), b = *reinterpret_cast<const float4*>(piv + row * kTcPs + c0 + 4);
		m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
	}
#pragma unroll
	for (int q = 0; q < 8; ++q)
	{
		const int p = 2 * q, ph = p >> 3, pc = p & 7;                      // half and registers (pc, pc + 1) holding columns p, p + 1
		const float b00 = __shfl_sync(kAll, m[pc], p + 16 * ph), b01 = __shfl_sync(kAll, m[pc + 1], p + 16 * ph);
		const float b10 = __shfl_sync(kAll, m[pc], p + 1 + 16 * ph), b11 = __shfl_sync(kAll, m[pc + 1], p + 1 + 16 * ph);
		const float f0 = __shfl_sync(kAll, m[pc], row + 16 * ph), f1 = __shfl_sync(kAll, m[pc + 1], row + 16 * ph);   // M[row][P]
		float r0[8], r1[8];
#pragma unroll
		for (int c = 0; c < 8; ++c)
		{
			r0[c] = __shfl_sync(kAll, m[c], p + 16 * half);                // M[p][my columns]
			r1[c] = __shfl_sync(kAll, m[c], p + 1 + 16 * half);            // M[p + 1][my columns]
		}
		const float rd = rcp_newton(__fmaf_rn(b00, b11, -__fmul_rn(b01, b10)));
		const float i00 = __fmul_rn(b11, rd), i01 = -__fmul_rn(b01, rd), i10 = -__fmul_rn(b10, rd), i11 = __fmul_rn(b00, rd);
		const bool isP0 = row == p, isP1 = row == p + 1, isP = isP0 || isP1;
		const float g0 = isP0 ? -i00 : (isP1 ? -i10 : __fmaf_rn(f0, i00, __fmul_rn(f1, i10)));
		const float g1 = isP0 ? -i01 : (isP1 ? -i11 : __fmaf_rn(f0, i01, __fmul_rn(f1, i11)));
#pragma unroll
		for (int c = 0; c < 8; ++c) m[c] = __fmaf_rn(-g1, r1[c], __fmaf_rn(-g0, r0[c], isP ? 0.0f : m[c]));
		if (half == ph) { m[pc] = -g0; m[pc + 1] = -g1; }
	}
	*reinterpret_cast<float4*>(P + row * kTcPs + c0) = make_float4(m[0], m[1], m[2], m[3]);
	*reinterpret_cast<float4*>(P + row * kTcPs + c0 + 4) = make_float4(m[4], m[5], m[6], m[7]);
	// P again as a tensor-core operand (B of the product C P^T: row n = this lane's row, k = c0 .. c0 + 7; P is symmetric)
	float hi[8], lo[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) tc::split_tf32(m[k], hi[k], lo[k]);
	unsigned char* h = reinterpret_cast<unsigned char*>(pHi) + tc::operand_offset(row, c0);
	unsigned char* l = reinterpret_cast<unsigned char*>(pLo) + tc::operand_offset(row, c0);
	*reinterpret_cast<float4*>(h) = make_float4(hi[0], hi[1], hi[2], hi[3]);
	*reinterpret_cast<float4*>(h + tc::kLbo) = make_float4(hi[4], hi[5], hi[6], hi[7]);
	*reinterpret_cast<float4*>(l) = make_float4(lo[0], lo[1], lo[2], lo[3]);
	*reinterpret_cast<float4*>(l + tc::kLbo) = make_float4(lo[4], lo[5], lo[6], lo[7]);
}

// one operand row (16 k) of thread / row `r`: hi and lo halves, four 16-byte stores each
__device__ __forceinline__ void store_operand_row(float* __restrict__ hiBuf, float* __restrict__ loBuf, const int r, const float (&v)[16])
{
	float hi[16], lo[16];
#pragma unroll
	for (int k = 0; k < 16; ++k) tc::split_tf32(v[k], hi[k], lo[k]);
	unsigned char* h = reinterpret_cast<unsigned char*>(hiBuf) + tc::operand_offset(r, 0);
	unsigned char* l = reinterpret_cast<unsigned char*>(loBuf) + tc::operand_offset(r, 0);
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		*reinterpret_cast<float4*>(h + q * tc::kLbo) = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
		*reinterpret_cast<float4*>(l + q * tc::kLbo) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
	}
}

// What the NEXT fine bank of this CTA will gather, so that the pivot warp (idle but for sixteen elimination steps per panel)
// can pull it into L2 while this system is being inverted: the gather is three dependent global round trips (s2o -> CSR
// range -> neighbour indices -> blocks; 10 k cycles per bank when they all go to HBM).  The chain is walked in stages, one
// per panel, each consuming what the previous stage loaded a few thousand cycles earlier: the warp never waits on memory.
struct TcPrefetch
{
	const int* s2o = nullptr;
	const int* adjStart = nullptr;
	const int* ranges = nullptr;
	const int* adjIdx = nullptr;
	const float* offdiag = nullptr;
	const float* diag = nullptr;
	int nv = 0, bank = -1;            // bank < 0: nothing to prefetch
};
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// In: s.A holds the assembled 96x96 system (row stride kLdP), all threads past the barrier that completed it.
// Out: s.packed holds the packed inverse (lane-slot layout, mas_internal.h), all threads past a barrier.
// pos96[r * 96 + c], r >= c: packed position of symmetric element (r, c).  `parity`: phase of s.bar, carried across systems.
__device__ __forceinline__ void invert_tile_tc(TcSmem& s, const TcAddr tb, uint32_t& parity, const unsigned short* __restrict__ pos96,
	int* __restrict__ errFlag, PhaseClock& pc, const TcPrefetch pf = TcPrefetch())
{
	// the warp index through a shuffle: the compiler then KNOWS it is warp-uniform and keeps the role branches below uniform
	// (from threadIdx alone it cannot, and wraps every shuffle inside `if (warp == 3)` in a convergence sequence)
	const int t = threadIdx.x, warp = __shfl_sync(0xffffffffu, t >> 5, 0), lane = t & 31;
	const uint32_t myLane = (uint32_t)(32 * warp) << 16;        // lane field of this warp's TMEM quadrant (the hardware adds the lane)

	// padding nodes: zero (0,0) entry of the diagonal block -> identity (cpp:1365-1368)
	if (t < kBank && s.A[tile_at(3 * t, 3 * t)] == 0.0f)
	{
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j) s.A[tile_at(3 * t + i, 3 * t + j)] = (i == j) ? 1.0f : 0.0f;
	}
	__syncthreads();

	// the system into tensor memory: thread r stores row r (conflict-free reads at the odd row stride)
	if (warp < 3)
	{
#pragma unroll
		for (int c0 = 0; c0 < kDof; c0 += 16)
		{
			float v[16];
#pragma unroll
			for (int j = 0; j < 16; ++j) v[j] = s.A[tile_at(t, c0 + j)];
			tc::tmem_st16(tb.col(c0) + myLane, v);
		}
		tc::tmem_wait_st();
	}
	tc::fence_before_sync();
	__syncthreads();                    // the tile array becomes the operand buffers
	tc::fence_after_sync();
	if (warp == 3)
	{
		// rows 96..127 of the A operand (M = 128 for the instruction, 96 rows of payload) are zero
		float4* zh = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s.op.aHi) + tc::operand_bytes(96));
		float4* zl = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s.op.aLo) + tc::operand_bytes(96));
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			zh[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
			zl[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
	}
	pc.mark(3);

	int pfOv = 0, pfE0 = 0, pfE1 = 0, pfSrc = 0;       // prefetch state of the pivot warp
#pragma unroll 1
	for (int K = 0; K < 6; ++K)
	{
		if (K > 0)
		{
			if (!tc::mbar_wait(&s.bar, parity) && t == 0) atomicExch(errFlag, 1);
			parity ^= 1u;
			tc::fence_after_sync();
		}
		pc.mark(4);
		const uint32_t colK = tb.col(16 * K);              // TMEM address of column block K (lane 0)
		const bool inK = (t >> 4) == K;                    // this thread's row belongs to block K
		float c[16];
		if (warp < 3)
		{
			tc::tmem_ld16(colK + myLane, c);               // C[r, :], and for the rows of block K the pivot block itself
			if (inK)
			{
				float4* dst = reinterpret_cast<float4*>(s.piv + (t & 15) * kTcPs);
#pragma unroll
				for (int q = 0; q < 4; ++q) dst[q] = make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
#pragma unroll
				for (int k = 0; k < 16; ++k) c[k] = 0.0f;  // Bop[x] = 0 for the rows of block K
			}
		}
		__syncthreads();
		pc.mark(5);
		if (warp < 3)
		{
			// (beside warp 3 inverting the pivot block)  Bop = C[r, :]: B operand of the update GEMM and A operand of C P
			store_operand_row(s.op.bHi, s.op.bLo, t, c);
			// row block K of T is REPLACED by this panel: zero it, the update GEMM then deposits P C^T there
			if (warp == (K >> 1))
			{
				tc::tmem_zero_16lanes_x8(tc::tmem_at(tb.a, 16 * K, 0));      // columns 0..63
				tc::tmem_zero_16lanes_x4(tc::tmem_at(tb.b, 16 * K, 0));      // columns 64..95
				tc::tmem_wait_st();
			}
		}
		if (warp == 3)
		{
			invert16_warp(s.piv, s.P, s.pHi, s.pLo, lane);
		}
		tc::fence_async_smem();
		tc::fence_before_sync();
		__syncthreads();
		pc.mark(6);
		const uint32_t aH = tc::smem_addr(s.op.aHi), aL = tc::smem_addr(s.op.aLo), bH = tc::smem_addr(s.op.bHi), bL = tc::smem_addr(s.op.bLo);
		if (t == 96)
		{
			// T[:, K] = C P^T (3xTF32, small terms first; the first MMA overwrites the column block)
			tc::fence_after_sync();
			constexpr uint32_t idP = tc::idesc_tf32(128, 16);
			const uint32_t pH = tc::smem_addr(s.pHi), pL = tc::smem_addr(s.pLo);
#pragma unroll
			for (int ks = 0; ks < 2; ++ks)
			{
				const uint32_t off = ks * 2 * tc::kLbo;
				tc::mma_tf32(colK, tc::smem_desc(bL + off, tc::kLbo, tc::kSbo), tc::smem_desc(pH + off, tc::kLbo, tc::kSbo), idP, ks);
				tc::mma_tf32(colK, tc::smem_desc(bH + off, tc::kLbo, tc::kSbo), tc::smem_desc(pL + off, tc::kLbo, tc::kSbo), idP, 1u);
				tc::mma_tf32(colK, tc::smem_desc(bH + off, tc::kLbo, tc::kSbo), tc::smem_desc(pH + off, tc::kLbo, tc::kSbo), idP, 1u);
			}
			tc::mma_commit(&s.bar);
		}
		if (warp == 3)
		{
			// staged prefetch of the next bank's inputs (lane = vertex), see TcPrefetch; here the pivot warp has nothing else to do
			// until the product C P^T completes
			const int vn = pf.bank * 32 + lane;
			if (pf.bank >= 0 && vn < pf.nv)
			{
				if (K == 0)
				{
					pfOv = pf.s2o[vn];
					pfE0 = pf.adjStart[vn];
					pfE1 = pf.adjStart[vn + 1];
				}
				else if (K == 1)
				{
					pfSrc = pf.ranges[pfOv];
					prefetch_l2(pf.adjIdx + pfE0);
					prefetch_l2(pf.adjIdx + pfE1 - 1);
					prefetch_l2(pf.diag + 9 * (size_t)pfOv);
				}
				else if (K == 2)
				{
					const char* first = reinterpret_cast<const char*>(pf.offdiag + 9 * (size_t)pfSrc);
					const int bytes = 36 * (pfE1 - pfE0);
					for (int off = 0; off < bytes + 127; off += 128) prefetch_l2(first + (off < bytes ? off : bytes - 1));
				}
			}
		}
		if (!tc::mbar_wait(&s.bar, parity) && t == 0) atomicExch(errFlag, 1);
		parity ^= 1u;
		tc::fence_after_sync();
		pc.mark(7);
		if (warp < 3)
		{
			float a[16];
			tc::tmem_ld16(colK + myLane, a);               // (C P)[r, :]; zero for the rows of block K
			if (warp == (K >> 1))
			{
				// the rows of block K store -P into the pivot block (their 16 lanes cannot be written alone with this shape:
				// the other 16 rows of the warp write back what they have just read)
				if (inK)
				{
					const float4* src = reinterpret_cast<const float4*>(s.P + (t & 15) * kTcPs);
#pragma unroll
					for (int q = 0; q < 4; ++q)
					{
						const float4 v = src[q];
						a[4 * q] = -v.x; a[4 * q + 1] = -v.y; a[4 * q + 2] = -v.z; a[4 * q + 3] = -v.w;
					}
				}
				tc::tmem_st16(colK + myLane, a);
				tc::tmem_wait_st();
			}
			// Aop = -(C P)[r, :] for the rows outside block K, P[x, :] for the rows of block K: the negative of `a` either way
#pragma unroll
			for (int k = 0; k < 16; ++k) a[k] = -a[k];
			store_operand_row(s.op.aHi, s.op.aLo, t, a);
		}
		tc::fence_async_smem();
		tc::fence_before_sync();
		__syncthreads();
		if (t == 96)
		{
			tc::fence_after_sync();
			constexpr uint32_t idA = tc::idesc_tf32(128, kTcColsA), idB = tc::idesc_tf32(128, kTcColsB);
			constexpr uint32_t rowsB = tc::operand_bytes(kTcColsA);       // B operand rows 64..95 feed matrix columns 64..95
#pragma unroll
			for (int ks = 0; ks < 2; ++ks)
			{
				const uint32_t off = ks * 2 * tc::kLbo;
				const uint64_t dAl = tc::smem_desc(aL + off, tc::kLbo, tc::kSbo), dAh = tc::smem_desc(aH + off, tc::kLbo, tc::kSbo);
				// small terms first: lo * hi, hi * lo, hi * hi; each as N = 64 (columns 0..63) and N = 32 (columns 64..95)
				tc::mma_tf32(tb.a, dAl, tc::smem_desc(bH + off, tc::kLbo, tc::kSbo), idA, 1u);
				tc::mma_tf32(tb.b, dAl, tc::smem_desc(bH + rowsB + off, tc::kLbo, tc::kSbo), idB, 1u);
				tc::mma_tf32(tb.a, dAh, tc::smem_desc(bL + off, tc::kLbo, tc::kSbo), idA, 1u);
				tc::mma_tf32(tb.b, dAh, tc::smem_desc(bL + rowsB + off, tc::kLbo, tc::kSbo), idB, 1u);
				tc::mma_tf32(tb.a, dAh, tc::smem_desc(bH + off, tc::kLbo, tc::kSbo), idA, 1u);
				tc::mma_tf32(tb.b, dAh, tc::smem_desc(bH + rowsB + off, tc::kLbo, tc::kSbo), idB, 1u);
			}
			tc::mma_commit(&s.bar);
		}
	}
	if (!tc::mbar_wait(&s.bar, parity) && t == 0) atomicExch(errFlag, 1);
	parity ^= 1u;
	tc::fence_after_sync();
	pc.mark(8);

	// T = -A^-1: lower triangle into the packed ("lane-slot") order
	if (warp < 3)
	{
#pragma unroll 1
		for (int c0 = 0; c0 <= 32 * warp + 16; c0 += 16)          // warp-uniform bound: columns up to the warp's last row
		{
			float v[16];
			const uint4 p0 = *reinterpret_cast<const uint4*>(pos96 + t * kDof + c0), p1 = *reinterpret_cast<const uint4*>(pos96 + t * kDof + c0 + 8);
			const unsigned pw[8] = { p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w };      // sixteen 16-bit positions
			tc::tmem_ld16(tb.col(c0) + myLane, v);
#pragma unroll
			for (int j = 0; j < 16; ++j)
				if (c0 + j <= t) s.packed[(pw[j >> 1] >> (16 * (j & 1))) & 0xffffu] = -v[j];
		}
	}
	tc::fence_before_sync();
	__syncthreads();
	pc.mark(9);
}

// per-CTA set-up and tear-down of the tensor-memory allocations and the mbarrier
__device__ __forceinline__ TcAddr tc_begin(TcSmem& s)
{
	if (threadIdx.x < 32)
	{
		tc::tmem_alloc_only<kTcColsA>(&s.tmemBase[0]);
		tc::tmem_alloc_only<kTcColsB>(&s.tmemBase[1]);
		tc::tmem_relinquish();
	}
	if (threadIdx.x == 32)
	{
		tc::mbar_init(&s.bar, 1);
		tc::mbar_init_fence();
	}
	tc::fence_before_sync();
	__syncthreads();
	tc::fence_after_sync();
	TcAddr tb;
	tb.a = s.tmemBase[0];
	tb.b = s.tmemBase[1];
	return tb;
}
__device__ __forceinline__ void tc_end(const TcAddr tb)
{
	tc::fence_before_sync();
	__syncthreads();
	if (threadIdx.x < 32)
	{
		tc::tmem_dealloc<kTcColsA>(tb.a);
		tc::tmem_dealloc<kTcColsB>(tb.b);
	}
}

// Dynamic distribution of the systems over the persistent CTAs: one counter per launch.  (A CTA that has to wait for tensor
// memory — several contexts can share a GPU — then simply takes fewer systems.)  Returns the next index, the same in every
// thread; contains a barrier.
__device__ __forceinline__ int tc_next_work(TcSmem& s, int* __restrict__ counter)
{
	if (threadIdx.x == 0) s.nextWork = atomicAdd(counter, 1);
	__syncthreads();
	return s.nextWork;
}
