// TEST INFRASTRUCTURE ONLY.  The apply kernels of csrc/mas_apply.cu (restrict_fine, restrict_l1, restrict_top, solve_coarse,
// prolong_sum, solve_fine — the text nvcc compiles, host launches guarded out) run on the CPU through tests/emu/cuda_emu.h,
// launched in the order of apply_begin + apply_end for a single-GPU context.
//   apply_emu < in.bin > out.bin
//   in : int32 nv, numLevel, totalClusters, levelSize[(numLevel + 1) * 2]; int32 s2o[nv]; int32 goingNext[totalClusters];
//        float32 dense inverses [totalClusters / 32][96][96]; float32 r[nv][4]; int32 coarseTables[nv][4]
//   env: MAS_EMU_TOP_FROM_L1 / MAS_EMU_WALK = the launch sequences apply_forked uses on small meshes
//   out: float32 z[nv][4]
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#define MAS_CPU_EMULATION 1
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_apply.cu"

template <typename T>
static std::vector<T> rd(size_t n)
{
	std::vector<T> v(n);
	if (n && fread(v.data(), sizeof(T), n, stdin) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

int main()
{
	using namespace mas;
	const auto hdr = rd<int>(3);
	const int nv = hdr[0], L = hdr[1], total = hdr[2];
	const auto ls = rd<int>((size_t)(L + 1) * 2);
	const auto s2o = rd<int>((size_t)nv);
	const auto goingNext = rd<int>((size_t)total);
	const int nBlocks = total / 32, nVC = pad32(nv), nFine = nVC / 32, nCoarse = total - nVC;
	std::vector<float> packed((size_t)nBlocks * kTri);
	{
		std::vector<float> dense((size_t)kDof * kDof);
		for (int b = 0; b < nBlocks; ++b)
		{
			if (fread(dense.data(), 4, dense.size(), stdin) != dense.size()) return 2;
			for (int r = 0; r < kDof; ++r)
				for (int c = 0; c <= r; ++c) packed[(size_t)b * kTri + packed_pos(r, c)] = dense[(size_t)r * kDof + c];
		}
	}
	const auto rIn = rd<float4>((size_t)nv);
	const auto coarseTables = rd<int4>((size_t)nv);
	const bool walk = getenv("MAS_EMU_WALK") != nullptr && L >= 2;    // head = all banks: level-0 solve first, coarse part added last
	std::vector<float4> z((size_t)nv, make_float4(7.f, 7.f, 7.f, 7.f));
	const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
	std::vector<float4> coarseR((size_t)(nCoarse > 0 ? nCoarse : 1), zero), coarseZ(coarseR), zsum(coarseR);
	const int top = L < 4 ? L : 4;                     // prolonged_top() without MAS_OPT_PROLONG_ALL_LEVELS

	if (walk)
		emu::launch(cdiv(nFine, kWarpsPerCta), kApplyThreads, [&] {
			solve_fine_kernel(packed.data(), rIn.data(), s2o.data(), goingNext.data(), zsum.data(), nv, nVC, 0, nFine, 0, 0, z.data());
		});
	if (L >= 2)
	{
		// apply_begin
		emu::launch(cdiv(nFine, kWarpsPerCta * kRestrictBanks), kApplyThreads, [&] {
			restrict_fine_kernel(rIn.data(), s2o.data(), goingNext.data(), nv, nVC, 0, nFine, coarseR.data(), nullptr, 0ull, nullptr);
		});
		// launch_coarse
		const int cnt1 = ls[2], begin1 = ls[3];
		const int nCoarseBlocks = nCoarse / 32, nL1Blocks = pad32(cnt1) / 32;
		const bool topFromL1 = getenv("MAS_EMU_TOP_FROM_L1") && L > 2 && cnt1 <= 512;     // small meshes
		bool fusedSolve = false;
		if (L > 2 && !topFromL1)
			emu::launch(cdiv(cdiv(cnt1, 32), kWarpsPerCta), kApplyThreads, [&] {
				restrict_l1_kernel(goingNext.data(), begin1, cnt1, nVC, 0, cdiv(cnt1, 32), coarseR.data(), nullptr, 0ull, nullptr);
			});
		if (L > 3 || topFromL1)
		{
			TopArgs a;
			a.numLevel = L; a.nVC = nVC; a.firstLevel = topFromL1 ? 1 : 2;
			for (int l = 0; l <= kMaxLevel; ++l) { a.count[l] = 0; a.begin[l] = 0; }
			for (int l = 1; l <= L; ++l) { a.count[l] = ls[2 * l]; a.begin[l] = ls[2 * l + 1]; }
			const int cnt2 = ls[4];
			if (cnt2 > 2048 && !topFromL1)
			{
				emu::launch(cdiv(cdiv(cnt2, 32), kWarpsPerCta), kApplyThreads, [&] {
					restrict_l1_kernel(goingNext.data(), ls[5], cnt2, nVC, 0, cdiv(cnt2, 32), coarseR.data(), nullptr, 0ull, nullptr);
				});
				a.firstLevel = 3;
			}
			if (topFromL1 && nCoarseBlocks <= kTopSolveBlocks)       // launch_coarse's one-CTA kernel for tiny hierarchies
			{
				emu::launch(1, kTopSolveThreads, [&] {
					top_solve_kernel(goingNext.data(), a, coarseR.data(), packed.data() + (size_t)nFine * kTri, coarseZ.data(), nCoarseBlocks);
				});
				fusedSolve = true;
			}
			else if (a.firstLevel + 1 < L)
				emu::launch(1, kTopThreads, [&] { restrict_top_kernel(goingNext.data(), a, coarseR.data()); });
		}
		if (nCoarseBlocks > 0 && !fusedSolve)
			emu::launch(nCoarseBlocks, 128, [&] {
				solve_coarse_kernel(packed.data() + (size_t)nFine * kTri, coarseR.data(), coarseZ.data(), 0, nL1Blocks, nL1Blocks);
			});
		if (cnt1 > 0 && !walk)
			emu::launch(cdiv(cnt1, 256), 256, [&] {
				prolong_sum_kernel(coarseZ.data(), goingNext.data(), begin1, 0, cnt1, nVC, top - 2, zsum.data());
			});
	}
	if (walk)
	{
		emu::launch(cdiv(nv, 256), 256, [&] {
			add_coarse_walk_kernel(s2o.data(), coarseTables.data(), coarseZ.data(), 0, nv, nVC, top - 1, z.data());
		});
		fwrite(z.data(), sizeof(float4), z.size(), stdout);
		return 0;
	}
	// apply_end
	emu::launch(cdiv(nFine, kWarpsPerCta), kApplyThreads, [&] {
		solve_fine_kernel(packed.data(), rIn.data(), s2o.data(), goingNext.data(), zsum.data(), nv, nVC, 0, nFine, 0, top >= 2 ? 1 : 0, z.data());
	});
	fwrite(z.data(), sizeof(float4), z.size(), stdout);
	return 0;
}
