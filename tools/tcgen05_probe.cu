// Development probe for csrc/mas_tcgen05.cuh: checks, on a real B200, the hand-written tcgen05 / TMEM primitives the batched
// inversion is built from, each against a CPU answer, and prints PASS / FAIL per item:
//   1. tcgen05.mma kind::tf32, M = 128, N = 96, K = 16 (two k-steps) from the K-major no-swizzle operand layout, accumulating
//      onto values written with tcgen05.st, read back with tcgen05.ld (lane = row, column = column), with the descriptor's
//      leading / stride byte offsets as documented and swapped;
//   2. the 16-lane zero store (tcgen05.st.16x256b) at lane offsets 0 and 16 of every warp quadrant;
//   3. four co-resident CTAs per SM each holding 128 TMEM columns (the occupancy the inversion kernel runs at);
//   4. latency of one ld -> barrier -> st -> fence -> barrier -> mma -> commit -> wait round trip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_tcgen05.cuh"

using namespace mas::tc;

constexpr int kM = 128, kN = 96, kK = 16;

struct ProbeSmem
{
	alignas(128) float aHi[kM * kK];
	alignas(128) float bHi[kN * kK];
	alignas(8) uint64_t bar;
	uint32_t tmemBase;
};

// out: [128][96] result of test 1; zeroOut: [8][128][96] after each 16-lane zero store; flags: status words
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, const float* C0, float* out, float* zeroOut, int* flags,
	uint32_t lbo, uint32_t sbo, long long* cycles)
{
	extern __shared__ __align__(128) unsigned char raw[];
	ProbeSmem& s = *reinterpret_cast<ProbeSmem*>(raw);
	const int t = threadIdx.x, warp = t >> 5;
	if (warp == 0) tmem_alloc<128>(&s.tmemBase);
	if (t == 0) { mbar_init(&s.bar, 1); mbar_init_fence(); }
	fence_before_sync();
	__syncthreads();
	fence_after_sync();
	const uint32_t tb = s.tmemBase;
	if (t == 0) flags[1] = (int)tb;

	// operands into the K-major core-matrix layout
	for (int i = t; i < kM * kK; i += 128) reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s.aHi) + operand_offset(i / kK, i % kK))[0] = A[i];
	for (int i = t; i < kN * kK; i += 128) reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s.bHi) + operand_offset(i / kK, i % kK))[0] = B[i];
	// initial accumulator: thread = row (lane), 96 columns, 16 at a time
	for (int c0 = 0; c0 < kN; c0 += 16)
	{
		float v[16];
		for (int j = 0; j < 16; ++j) v[j] = C0[t * kN + c0 + j];
		tmem_st16(tmem_at(tb, 32 * warp, c0), v);
	}
	tmem_wait_st();
	fence_async_smem();
	fence_before_sync();
	__syncthreads();
	if (t == 96)
	{
		fence_after_sync();
		const uint32_t id = idesc_tf32(kM, kN);
		for (int ks = 0; ks < 2; ++ks)
			mma_tf32(tb, smem_desc(smem_addr(s.aHi) + ks * 2 * kLbo, lbo, sbo), smem_desc(smem_addr(s.bHi) + ks * 2 * kLbo, lbo, sbo), id, 1u);
		mma_commit(&s.bar);
	}
	const bool ok = mbar_wait(&s.bar, 0);
	if (!ok && t == 0) flags[0] = 1;
	fence_after_sync();
	for (int c0 = 0; c0 < kN; c0 += 16)
	{
		float v[16];
		tmem_ld16(tmem_at(tb, 32 * warp, c0), v);
		for (int j = 0; j < 16; ++j) out[t * kN + c0 + j] = v[j];
	}
	__syncthreads();

	// test 2: zero 16 lanes at a time
	for (int blk = 0; blk < 8; ++blk)
	{
		// refill with a marker
		for (int c0 = 0; c0 < kN; c0 += 16)
		{
			float v[16];
			for (int j = 0; j < 16; ++j) v[j] = 1.0f + t;
			tmem_st16(tmem_at(tb, 32 * warp, c0), v);
		}
		tmem_wait_st();
		fence_before_sync();
		__syncthreads();
		fence_after_sync();
		if (warp == blk / 2)
		{
			tmem_zero_16lanes_x8(tmem_at(tb, 16 * blk, 0));
			tmem_zero_16lanes_x4(tmem_at(tb, 16 * blk, 64));
			tmem_wait_st();
		}
		fence_before_sync();
		__syncthreads();
		fence_after_sync();
		for (int c0 = 0; c0 < kN; c0 += 16)
		{
			float v[16];
			tmem_ld16(tmem_at(tb, 32 * warp, c0), v);
			for (int j = 0; j < 16; ++j) zeroOut[((size_t)blk * 128 + t) * kN + c0 + j] = v[j];
		}
		__syncthreads();
	}

	// test 4: round-trip latency of the per-panel protocol (no arithmetic), 64 rounds
	uint32_t parity = 1;
	long long t0 = clock64();
	for (int round = 0; round < 64; ++round)
	{
		float v[16];
		tmem_ld16(tmem_at(tb, 32 * warp, 16 * (round % 6)), v);
		__syncthreads();
		tmem_st16(tmem_at(tb, 32 * warp, 16 * (round % 6)), v);
		tmem_wait_st();
		fence_async_smem();
		fence_before_sync();
		__syncthreads();
		if (t == 96)
		{
			fence_after_sync();
			const uint32_t id = idesc_tf32(kM, kN);
			for (int rep = 0; rep < 3; ++rep)
				for (int ks = 0; ks < 2; ++ks)
					mma_tf32(tb, smem_desc(smem_addr(s.aHi) + ks * 2 * kLbo, lbo, sbo), smem_desc(smem_addr(s.bHi) + ks * 2 * kLbo, lbo, sbo), id, 1u);
			mma_commit(&s.bar);
		}
		if (!mbar_wait(&s.bar, parity) && t == 0) flags[0] = 2;
		parity ^= 1;
		fence_after_sync();
	}
	if (t == 0 && blockIdx.x == 0) cycles[0] = (clock64() - t0) / 64;
	__syncthreads();
	if (warp == 0) tmem_dealloc<128>(tb);
}

int main()
{
	std::vector<float> A(kM * kK), B(kN * kK), C0(kM * kN), ref(kM * kN);
	srand(7);
	for (auto& v : A) v = (float)(rand() % 17 - 8);
	for (auto& v : B) v = (float)(rand() % 13 - 6);
	for (auto& v : C0) v = (float)(rand() % 101 - 50);
	for (int i = 0; i < kM; ++i)
		for (int j = 0; j < kN; ++j)
		{
			double acc = C0[i * kN + j];
			for (int k = 0; k < kK; ++k) acc += (double)A[i * kK + k] * B[j * kK + k];
			ref[i * kN + j] = (float)acc;
		}
	float *dA, *dB, *dC, *dOut, *dZero;
	int* dFlags;
	long long* dCyc;
	cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C0.size() * 4);
	cudaMalloc(&dOut, ref.size() * 4); cudaMalloc(&dZero, 8 * ref.size() * 4); cudaMalloc(&dFlags, 16); cudaMalloc(&dCyc, 8);
	cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(dC, C0.data(), C0.size() * 4, cudaMemcpyHostToDevice);
	const int smem = 48 * 1024;       // as in the inversion kernel: at most four CTAs per SM
	cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	const uint32_t cand[2][2] = { { kLbo, kSbo }, { kSbo, kLbo } };
	int rc = 0;
	for (int ci = 0; ci < 2; ++ci)
	{
		cudaMemset(dFlags, 0, 16); cudaMemset(dOut, 0, ref.size() * 4); cudaMemset(dZero, 0xff, 8 * ref.size() * 4);
		const int grid = ci == 0 ? 148 * 4 : 1;       // candidate 0 also exercises four CTAs per SM on every SM
		probe_kernel<<<grid, 128, smem>>>(dA, dB, dC, dOut, dZero, dFlags, cand[ci][0], cand[ci][1], dCyc);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("candidate %d: CUDA error %s\n", ci, cudaGetErrorString(e)); return 2; }
		std::vector<float> out(ref.size()), zo(8 * ref.size());
		int flags[4];
		long long cyc = 0;
		cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(zo.data(), dZero, zo.size() * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(flags, dFlags, 16, cudaMemcpyDeviceToHost);
		cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost);
		int bad = 0, badUsed = 0;
		for (int i = 0; i < kM * kN; ++i)
		{
			if (out[i] != ref[i]) { ++bad; if (i / kN < 96) ++badUsed; }
		}
		printf("candidate %d (lbo %u, sbo %u): mma %s (%d of %d mismatches, %d in rows < 96), timeout flag %d, tmem base 0x%x\n", ci, cand[ci][0], cand[ci][1],
			bad == 0 ? "PASS" : "FAIL", bad, kM * kN, badUsed, flags[0], flags[1]);
		if (bad) { printf("  first rows: out[0][0..3] = %g %g %g %g, ref = %g %g %g %g\n", out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]); }
		if (ci == 0)
		{
			if (bad) rc = 1;
			for (int blk = 0; blk < 8; ++blk)
			{
				int wrong = 0;
				for (int r = 0; r < 128; ++r)
					for (int c = 0; c < kN; ++c)
					{
						const float want = (r >= 16 * blk && r < 16 * blk + 16) ? 0.0f : 1.0f + r;
						if (zo[((size_t)blk * 128 + r) * kN + c] != want) ++wrong;
					}
				printf("  16-lane zero store, lanes %3d..%3d: %s (%d wrong)\n", 16 * blk, 16 * blk + 15, wrong == 0 ? "PASS" : "FAIL", wrong);
				if (wrong && blk < 6) rc = 1;
			}
			printf("  per-panel protocol round trip (ld, 2 barriers, st, 6 MMAs, commit, wait): %lld cycles; grid %d CTAs of 128 threads, 48 KB smem, 128 TMEM columns each\n", cyc, grid);
		}
	}
	return rc;
}
