// Blackwell (sm_100a) tensor-core primitives used by the batched inversion (mas_invert_tc.cuh): tcgen05.mma kind::tf32 with
// shared-memory operand descriptors, tensor-memory (TMEM) allocation / load / store, mbarrier completion.  Inline PTX only.
//
// Field layouts follow the PTX ISA's "matrix descriptor" and "instruction descriptor" tables (the same bit positions CUTLASS
// spells out in cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mas {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- operand layout in shared memory: K-major, no swizzle ("interleave") ------------------------------------------------
// A tf32 operand of R rows x 16 k is stored as 8-row x 16-byte "core matrices" (8 rows x 4 k, 128 contiguous bytes):
//     element (r, k)  ->  byte  (r >> 3) * kSbo + (r & 7) * 16 + (k >> 2) * kLbo + (k & 3) * 4
// kLbo: distance between core matrices that are adjacent in k; kSbo: distance between 8-row groups.
// One tcgen05.mma kind::tf32 consumes k = 8 (two core matrices per row group); the second k-step starts 2 * kLbo further.
constexpr uint32_t kLbo = 128;
constexpr uint32_t kSbo = 512;     // four core matrices (k = 16) per 8-row group
__host__ __device__ constexpr uint32_t operand_offset(int r, int k)
{
	return (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16u + (uint32_t)(k >> 2) * kLbo + (uint32_t)(k & 3) * 4u;
}
__host__ __device__ constexpr uint32_t operand_bytes(int rows) { return (uint32_t)(rows >> 3) * kSbo; }

// shared-memory matrix descriptor: start address [0,14) >> 4, leading byte offset [16,30) >> 4, stride byte offset [32,46) >> 4,
// descriptor version 1 at [46,48) (Blackwell), base offset 0, swizzle mode [61,64) = 0 (none)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes)
{
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3fffu);
	d |= (uint64_t)((lboBytes >> 4) & 0x3fffu) << 16;
	d |= (uint64_t)((sboBytes >> 4) & 0x3fffu) << 32;
	d |= (uint64_t)1 << 46;
	return d;
}

// instruction descriptor, kind::tf32, FP32 accumulate, A and B K-major, dense: c_format [4,6) = 1 (F32), a_format [7,10) = 2
// (TF32), b_format [10,13) = 2, a_major bit 15 = 0, b_major bit 16 = 0, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
		"}\n" ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory made visible to the async proxy (the tensor core reads its operands through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Waits for the phase with the given parity; gives up after ~2^26 polls (a descriptor or protocol bug must not hang the GPU)
// and returns false.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity)
{
	const uint32_t a = smem_addr(bar);
#pragma unroll 1
	for (uint32_t spin = 0; spin < (1u << 26); ++spin)
	{
		uint32_t done;
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

// ---- tensor memory ------------------------------------------------------------------------------------------------------
// address = (lane << 16) | column.  A warp reaches the 32 lanes of its own quadrant, lanes 32 * (warp % 4) ... + 31.
__device__ __forceinline__ uint32_t tmem_at(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }

// one warp: allocate nCols (power of two >= 32) columns, address lands in *slot (shared memory); after the CTA's last
// allocation tmem_relinquish() lets the other CTAs of the SM allocate without waiting for this one to exit
template <int nCols>
__device__ __forceinline__ void tmem_alloc_only(uint32_t* slot)
{
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "n"(nCols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
template <int nCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot)
{
	tmem_alloc_only<nCols>(slot);
	tmem_relinquish();
}
template <int nCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t base)
{
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(nCols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, 16 consecutive columns: thread t of the warp gets lane (quadrant base + t), columns col .. col + 15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
	uint32_t r[16];
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
		  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		: "r"(taddr)
		: "memory");
	tmem_wait_ld();
#pragma unroll
	for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16])
{
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
		"r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
		"r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
		"r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
		"r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
		: "memory");
}
// 16 lanes x 256 bit pattern: the warp writes 16 lanes (lane field of taddr = first lane, a multiple of 16 inside the warp's
// quadrant) x 8 columns per repetition; used only to ZERO a 16-row block, so the register-to-element map does not matter.
__device__ __forceinline__ void tmem_zero_16lanes_x8(uint32_t taddr)   // 16 lanes x 64 columns
{
	const uint32_t z = 0u;
	asm volatile(
		"tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
		"%1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
		"r"(z)
		: "memory");
}
__device__ __forceinline__ void tmem_zero_16lanes_x4(uint32_t taddr)   // 16 lanes x 32 columns
{
	const uint32_t z = 0u;
	asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
		"r"(z)
		: "memory");
}

// ---- 3xTF32 operand split ------------------------------------------------------------------------------------------------
// x = hi + lo with both halves exactly representable in TF32 (10 explicit mantissa bits); hi*hi + hi*lo + lo*hi recovers an
// FP32-accurate product on the TF32 tensor-core path (tools/sweep_inversion_study.py).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo)
{
	uint32_t h, l;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
	hi = __uint_as_float(h);
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(__fsub_rn(x, hi)));
	lo = __uint_as_float(l);
}

}  // namespace tc
}  // namespace mas
