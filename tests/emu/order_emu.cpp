// TEST INFRASTRUCTURE ONLY.  The ordering kernels of csrc/mas_order.cu — aabb_partial (two stages), morton, inverse_perm,
// sorted_degree, remap_adjacency — run on the CPU through tests/emu/cuda_emu.h in the order of order_vertices(); the CUB
// radix sort (stable, key = 63-bit code, payload = original index) and the CUB scan are replaced by std::stable_sort and a
// loop, which is what they compute.
//   order_emu < in.bin > out.bin
//   in : int32 nv, nnz; float32 positions[nv][4]; int32 starts[nv + 1], idx[nnz]   (caller's arrays, original order)
//   out: float32 aabb[8]; uint64 code[nv]; int32 s2o[nv], o2s[nv], adjStart[nv + 1], adjIdx[nnz]
#include "cuda_emu.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#define MAS_CPU_EMULATION 1
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_order.cu"

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
	const auto hdr = rd<int>(2);
	const int nv = hdr[0], nnz = hdr[1], threads = 256;
	const auto pos = rd<float4>((size_t)nv);
	const auto inStarts = rd<int>((size_t)nv + 1);
	const auto inIdx = rd<int>((size_t)nnz);

	int nPartials = 148 * 4;
	if (nPartials > cdiv(nv, kReduceThreads)) nPartials = cdiv(nv, kReduceThreads);
	if (nPartials < 1) nPartials = 1;
	std::vector<float4> partials((size_t)2 * nPartials), aabb(2);
	emu::launch(nPartials, kReduceThreads, [&] { aabb_partial_kernel(pos.data(), pos.data(), nv, partials.data(), nPartials); });
	emu::launch(1, kReduceThreads, [&] { aabb_partial_kernel(partials.data(), partials.data() + nPartials, nPartials, aabb.data(), 1); });

	std::vector<unsigned long long> code((size_t)nv);
	std::vector<int> iota((size_t)nv + 1), s2o((size_t)nv), o2s((size_t)nv);
	emu::launch(cdiv(nv, threads), threads, [&] { morton_kernel(pos.data(), aabb.data(), nv, code.data(), iota.data()); });
	std::copy(iota.begin(), iota.begin() + nv, s2o.begin());
	std::stable_sort(s2o.begin(), s2o.end(), [&](int a, int b) { return code[a] < code[b]; });      // cub::DeviceRadixSort::SortPairs
	emu::launch(cdiv(nv, threads), threads, [&] { inverse_perm_kernel(s2o.data(), nv, o2s.data()); });

	std::vector<int> deg((size_t)nv + 1), adjStart((size_t)nv + 1), adjIdx((size_t)(nnz > 0 ? nnz : 1));
	emu::launch(cdiv(nv + 1, threads), threads, [&] { sorted_degree_kernel(s2o.data(), inStarts.data(), nv, deg.data()); });
	std::exclusive_scan(deg.begin(), deg.end(), adjStart.begin(), 0);                                 // cub::DeviceScan::ExclusiveSum
	emu::launch(cdiv(nv, threads), threads, [&] {
		remap_adjacency_kernel(s2o.data(), o2s.data(), inStarts.data(), inIdx.data(), adjStart.data(), nv, adjIdx.data());
	});

	fwrite(aabb.data(), 16, 2, stdout);
	fwrite(code.data(), 8, (size_t)nv, stdout);
	fwrite(s2o.data(), 4, (size_t)nv, stdout);
	fwrite(o2s.data(), 4, (size_t)nv, stdout);
	fwrite(adjStart.data(), 4, (size_t)nv + 1, stdout);
	fwrite(adjIdx.data(), 4, (size_t)nnz, stdout);
	return 0;
}
