// TEST INFRASTRUCTURE ONLY.  The clustering kernels of csrc/mas_cluster.cu — connect_mask_l0 / _lx, collision_connect,
// close_components, exclusive_scan, number_components, next_level_table, coarse_tables — run on the CPU through
// tests/emu/cuda_emu.h, launched in the order of build_hierarchy + number_level for a single-GPU context.
//   cluster_emu < in.bin > out.bin
//   in : int32 nv, nnz, nStencil; int32 adjStart[nv + 1], adjIdx[nnz] (sorted space); nStencil 80-byte Stencil records and
//        int32 stencilIndexMapped[nStencil][5]
//   out: int32 numLevel, totalClusters, levelSize[(numLevel + 1) * 2], goingNext[totalClusters]; uint32 fineMask[nv];
//        int32 coarseSpaceTables[numLevel][nv], coarseTables[nv][4]
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#define MAS_CPU_EMULATION 1
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_cluster.cu"

template <typename T>
static std::vector<T> rd(size_t n)
{
	std::vector<T> v(n ? n : 1);
	if (n && fread(v.data(), sizeof(T), n, stdin) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

using namespace mas;

static std::vector<int> goingNext;

// number_level() of csrc/mas_cluster.cu
static int number_level(unsigned* mask, int count, int addSelf, int begin, int* idOut)
{
	const int threads = 256, nBanks = (count + 31) / 32;
	std::vector<int> bankCount((size_t)nBanks + 1), bankPrefix((size_t)nBanks + 1);
	int total = 0;
	emu::launch(cdiv((long long)nBanks * 32, threads), threads, [&] { close_components_kernel(mask, count, addSelf, bankCount.data()); });
	emu::launch(1, kScanThreads, [&] { exclusive_scan_kernel(bankCount.data(), nBanks, bankPrefix.data(), &total); });
	const int nextBegin = begin + pad32(count);
	if ((size_t)nextBegin > goingNext.size()) goingNext.resize((size_t)nextBegin * 2, 0);
	emu::launch(cdiv((long long)nBanks * 32, threads), threads, [&] {
		number_components_kernel(mask, count, bankPrefix.data(), begin, nextBegin, idOut, goingNext.data());
	});
	return total;
}

int main()
{
	const auto hdr = rd<int>(3);
	const int nv = hdr[0], nnz = hdr[1], nStencil = hdr[2], threads = 256;
	const auto adjStart = rd<int>((size_t)nv + 1);
	const auto adjIdx = rd<int>((size_t)nnz);
	const auto stencils = rd<Stencil>((size_t)nStencil);
	const auto stIdx = rd<int>((size_t)nStencil * 5);
	const int nVC = pad32(nv);
	int L = 1;
	for (int sz = nVC; sz > 32;) { sz /= 32; ++L; sz = pad32(sz); }          // level_count() of csrc/mas_api.cu
	std::vector<int> levelSize((size_t)(L + 2) * 2, 0);
	std::vector<unsigned> fineMask((size_t)nVC, 0u);
	std::vector<std::vector<int>> cst((size_t)L, std::vector<int>((size_t)nv, 0));
	goingNext.assign((size_t)nVC + (size_t)nVC / 8 + 4096, 0);

	emu::launch(cdiv(nVC, threads), threads, [&] { connect_mask_l0_kernel(adjStart.data(), adjIdx.data(), nv, nVC, fineMask.data()); });
	if (nStencil > 0)
		emu::launch(cdiv(nStencil, threads), threads, [&] {
			collision_connect_kernel(stencils.data(), stIdx.data(), nStencil, nullptr, fineMask.data());
		});
	const int n1 = number_level(fineMask.data(), nv, 0, 0, cst[0].data());
	levelSize[2] = n1; levelSize[3] = nVC;
	for (int level = 1; level < L; ++level)
	{
		const int cnt = levelSize[2 * level], begin = levelSize[2 * level + 1];
		std::vector<unsigned> nextMask((size_t)pad32(cnt) + 32, 0u);
		std::vector<int> nextId((size_t)pad32(cnt) + 32, 0);
		emu::launch(cdiv(pad32(nv), threads), threads, [&] {
			connect_mask_lx_kernel(adjStart.data(), adjIdx.data(), cst[level - 1].data(), nv, nextMask.data());
		});
		if (nStencil > 0)
			emu::launch(cdiv(nStencil, threads), threads, [&] {
				collision_connect_kernel(stencils.data(), stIdx.data(), nStencil, cst[level - 1].data(), nextMask.data());
			});
		const int nNext = number_level(nextMask.data(), cnt, 1, begin, nextId.data());
		levelSize[2 * (level + 1)] = nNext;
		levelSize[2 * (level + 1) + 1] = begin + pad32(cnt);
		emu::launch(cdiv(nv, threads), threads, [&] { next_level_table_kernel(cst[level - 1].data(), nextId.data(), nv, cst[level].data()); });
	}
	const int total = levelSize[2 * L + 1];
	if ((size_t)total > goingNext.size()) goingNext.resize((size_t)total, 0);
	std::vector<int4> coarseTables((size_t)nv);
	emu::launch(cdiv(nv, threads), threads, [&] { coarse_tables_kernel(goingNext.data(), nv, L, coarseTables.data()); });

	fwrite(&L, 4, 1, stdout);
	fwrite(&total, 4, 1, stdout);
	fwrite(levelSize.data(), 4, (size_t)(L + 1) * 2, stdout);
	fwrite(goingNext.data(), 4, (size_t)total, stdout);
	fwrite(fineMask.data(), 4, (size_t)nv, stdout);
	for (int l = 0; l < L; ++l) fwrite(cst[l].data(), 4, (size_t)nv, stdout);
	fwrite(coarseTables.data(), 16, (size_t)nv, stdout);
	return 0;
}
