// TEST INFRASTRUCTURE ONLY.  The stencil kernels of csrc/mas_cluster.cu — stencil_flag, exclusive_scan, stencil_build — run on
// the CPU through tests/emu/cuda_emu.h in the order of build_stencils() (literal Q2/Q3 reading or MAS_OPT_STENCIL_FIX).
//   stencil_emu < in.bin > out.bin
//   in : int32 nv, ne, nf, efN, eeN, vfN, fix, efRecords, eeRecords, vfRecords; int32 edges[ne][4], faces[nf][4], o2s[nv];
//        48-byte EfSet[efRecords], EeSet[eeRecords], VfSet[vfRecords]
//   out: int32 count; count 80-byte Stencil records; int32 stencilIndexMapped[count][5]
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

int main()
{
	using namespace mas;
	const auto h = rd<int>(10);
	const int nv = h[0], ne = h[1], nf = h[2], efN = h[3], eeN = h[4], vfN = h[5], fix = h[6];
	const auto edges = rd<int4>((size_t)ne);
	const auto faces = rd<int4>((size_t)nf);
	const auto o2s = rd<int>((size_t)nv);
	const auto ef = rd<unsigned char>((size_t)h[7] * 48);
	const auto ee = rd<unsigned char>((size_t)h[8] * 48);
	const auto vf = rd<unsigned char>((size_t)h[9] * 48);
	long long total = (long long)efN + eeN + vfN;
	const long long cap = (long long)nv * kMaxCollisionPerVert;
	if (total > cap) total = cap;
	const int n = (int)total, threads = 256;
	int count = 0, inputErr = 0;
	std::vector<int> flag((size_t)n + 1), slot((size_t)n + 1);
	if (n > 0)
	{
		emu::launch(cdiv(n, threads), threads, [&] { stencil_flag_kernel(ef.data(), ee.data(), vf.data(), efN, eeN, n, fix, nv, ne, nf, flag.data(), &inputErr); });
		emu::launch(1, kScanThreads, [&] { exclusive_scan_kernel(flag.data(), n, slot.data(), &count); });
	}
	std::vector<Stencil> out((size_t)(count > 0 ? count : 1));
	std::vector<int> outIdx((size_t)(count > 0 ? count : 1) * 5);
	if (count > 0)
		emu::launch(cdiv(n, threads), threads, [&] {
			stencil_build_kernel(ef.data(), ee.data(), vf.data(), efN, eeN, n, fix, flag.data(), slot.data(), edges.data(), faces.data(), o2s.data(),
				out.data(), outIdx.data());
		});
	fwrite(&count, 4, 1, stdout);
	fwrite(out.data(), sizeof(Stencil), (size_t)count, stdout);
	fwrite(outIdx.data(), 4, (size_t)count * 5, stdout);
	return 0;
}
