// TEST INFRASTRUCTURE ONLY.  The setup kernels of csrc/mas_assemble.cu — cross_bank, fine_assemble_invert, carry_up,
// coarse_invert (and through them csrc/mas_invert.cuh) — run on the CPU through tests/emu/cuda_emu.h, launched in the order of
// assemble_and_invert_begin + _end for a single-GPU context (collision_hessian included when stencils are given).
//   assemble_emu < in.bin > out.bin
//   in : int32 nv, numLevel, totalClusters, nnz, levelSize[(numLevel + 1) * 2]; int32 s2o[nv], adjStart[nv + 1], adjIdx[nnz],
//        goingNext[totalClusters], ranges[nv + 1]; float32 diag[nv][9], offdiag[nnz][9]  (caller's arrays, original order);
//        int32 nStencil; then nStencil 80-byte Stencil records and int32 stencilIndexMapped[nStencil][5]  (may be 0)
//   out: float32 dense inverses [totalClusters / 32][96][96]
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#define MAS_CPU_EMULATION 1
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_assemble.cu"

template <typename T>
static std::vector<T> rd(size_t n)
{
	std::vector<T> v(n ? n : 1);
	if (n && fread(v.data(), sizeof(T), n, stdin) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

int main()
{
	using namespace mas;
	const auto hdr = rd<int>(4);
	const int nv = hdr[0], L = hdr[1], total = hdr[2], nnz = hdr[3];
	const auto ls = rd<int>((size_t)(L + 1) * 2);
	const auto s2o = rd<int>((size_t)nv);
	const auto adjStart = rd<int>((size_t)nv + 1);
	const auto adjIdx = rd<int>((size_t)nnz);
	const auto goingNext = rd<int>((size_t)total);
	const auto ranges = rd<int>((size_t)nv + 1);
	const auto diag = rd<float>((size_t)nv * 9);
	const auto offdiag = rd<float>((size_t)nnz * 9);
	const int nVC = pad32(nv), nFine = nVC / 32, nCoarseNodes = total - nVC, nCoarseBlocks = nCoarseNodes / 32, nBlocks = total / 32;

	std::vector<double> acc((size_t)nCoarseBlocks * kDof * kDof + (size_t)nCoarseNodes * 9 + 1, 0.0);
	double* dense = acc.data();
	double* carry = acc.data() + (size_t)nCoarseBlocks * kDof * kDof;
	std::vector<float> packed((size_t)nBlocks * kTri, 0.f);
	std::vector<unsigned short> posTab((size_t)kOutPerThread * kInvThreads);   // ensure_pos_table()
	for (int t = 0; t < kInvThreads; ++t)
	{
		const int tr = t & 15, tc = t >> 4;
		int e = 0;
		for (int i = 0; i < 6; ++i)
			for (int j = 0; j < i; ++j, ++e) posTab[(size_t)e * kInvThreads + t] = (unsigned short)packed_pos(tr + 16 * i, tc + 16 * j);
		for (int i = 0; i < 6; ++i)
			posTab[(size_t)(15 + i) * kInvThreads + t] = (unsigned short)(tr >= tc ? packed_pos(tr + 16 * i, tc + 16 * i) : 0);
	}

	const int nStencil = rd<int>(1)[0];
	const auto stencils = rd<Stencil>((size_t)nStencil);
	const auto stIdx = rd<int>((size_t)nStencil * 5);
	std::vector<float> extraFine, cooVal;
	std::vector<int> cooCount, cooStart, cooFill;
	if (nStencil > 0)
	{
		extraFine.assign((size_t)nv * 9, 0.f);
		cooCount.assign((size_t)nFine, 0); cooStart.assign((size_t)nFine, 0); cooFill.assign((size_t)nFine, 0);
		CollisionArgs ca;
		ca.st = stencils.data(); ca.stIdx = stIdx.data(); ca.nStencil = nStencil;
		ca.goingNext = goingNext.data(); ca.numLevel = L; ca.nVC = nVC;
		ca.ownBegin = 0; ca.ownEnd = nVC;
		ca.extraFine = extraFine.data(); ca.dense = dense; ca.carry = carry;
		ca.cooCount = cooCount.data(); ca.cooStart = nullptr; ca.cooFill = cooFill.data(); ca.cooVal = nullptr;
		emu::launch(cdiv(nStencil, 256), 256, [&] { collision_hessian_kernel(ca, 0); });
		int entries = 0;                                 // launch_exclusive_scan
		for (int b = 0; b < nFine; ++b) { cooStart[b] = entries; entries += cooCount[b]; }
		cooVal.assign((size_t)(entries > 0 ? entries : 1) * 10, 0.f);
		ca.cooStart = cooStart.data(); ca.cooVal = cooVal.data();
		emu::launch(cdiv(nStencil, 256), 256, [&] { collision_hessian_kernel(ca, 1); });
	}

	FineArgs fa;
	fa.diag = diag.data(); fa.offdiag = offdiag.data(); fa.ranges = ranges.data();
	fa.s2o = s2o.data(); fa.adjStart = adjStart.data(); fa.adjIdx = adjIdx.data(); fa.goingNext = goingNext.data();
	fa.extraFine = nStencil > 0 ? extraFine.data() : nullptr;
	fa.cooStart = nStencil > 0 ? cooStart.data() : nullptr;
	fa.cooCount = nStencil > 0 ? cooCount.data() : nullptr;
	fa.cooVal = nStencil > 0 ? cooVal.data() : nullptr;
	fa.dense = dense; fa.carry = carry;
	fa.packedOut = packed.data();
	fa.posTab = posTab.data();
	fa.nv = nv; fa.nVC = nVC; fa.numLevel = L; fa.bankBegin = 0;

	// assemble_and_invert_begin
	if (L > 1) emu::launch(cdiv(nVC, 256), 256, [&] { if (L >= 5) cross_bank_kernel<true>(fa, 0, nVC); else cross_bank_kernel<false>(fa, 0, nVC); });
	emu::launch(nFine, kInvThreads, [&] { fine_assemble_invert_kernel(fa); });
	// assemble_and_invert_end
	for (int level = 1; level + 1 < L; ++level)
	{
		const int cnt = ls[2 * level], begin = ls[2 * level + 1];
		if (cnt <= 0) continue;
		emu::launch(cdiv((long long)cnt * 9, 256), 256, [&] { carry_up_kernel(carry, goingNext.data(), begin, cnt, nVC); });
	}
	if (nCoarseBlocks > 0)
	{
		const int nL1Blocks = pad32(ls[2]) / 32;
		emu::launch(nCoarseBlocks, kInvThreads, [&] {
			coarse_invert_kernel(dense, carry, packed.data() + (size_t)nFine * kTri, posTab.data(), 0, nL1Blocks, nL1Blocks);
		});
	}
	std::vector<float> out((size_t)kDof * kDof);
	for (int b = 0; b < nBlocks; ++b)
	{
		for (int r = 0; r < kDof; ++r)
			for (int c = 0; c < kDof; ++c) out[(size_t)r * kDof + c] = packed[(size_t)b * kTri + packed_pos(r, c)];
		fwrite(out.data(), 4, out.size(), stdout);
	}
	return 0;
}
