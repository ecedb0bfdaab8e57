// Drop-in check: a caller written against the reference's class API (SeSchwarzPreconditioner.h:44-63), compiled with
// g++ against include/SeSchwarzPreconditioner.h and linked to libmas_b200.so.  Reads a mesh dumped by the Python test,
// runs AllocatePrecoditioner / PreparePreconditioner / Preconditioning on host pointers, writes z.
//   dropin_main <in.bin> <out.bin>
// in.bin : int32 nv, ne, nf, nnz, efTotal, eeTotal, vfTotal, nStencilRecords; then positions[nv*4] f32, edges[ne*4] i32,
//          faces[nf*4] i32, starts[nv+1] i32, idx[nnz] i32, diag[nv*9] f32, offdiag[nnz*9] f32,
//          ef/ee/vf [nStencilRecords*48] bytes each, residual[nv*4] f32
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "SeSchwarzPreconditioner.h"

static_assert(offsetof(SE::EfSet, m_bary) == 12 && offsetof(SE::EfSet, m_normal) == 32, "EfSet layout");
static_assert(offsetof(SE::EeSet, m_bary) == 16 && offsetof(SE::EeSet, m_normal) == 32, "EeSet layout");
static_assert(offsetof(SE::VfSet, m_bary) == 16 && offsetof(SE::VfSet, m_normal) == 32, "VfSet layout");

template <typename T>
static std::vector<T> rd(FILE* f, size_t n)
{
	std::vector<T> v(n);
	if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

int main(int argc, char** argv)
{
	if (argc < 3) return 2;
	FILE* f = fopen(argv[1], "rb");
	if (!f) return 2;
	auto hdr = rd<int>(f, 8);
	const int nv = hdr[0], ne = hdr[1], nf = hdr[2], nnz = hdr[3], nRec = hdr[7];
	auto pos = rd<SE::SeVec3fSimd>(f, nv);
	auto edges = rd<SE::Int4>(f, ne);
	auto faces = rd<SE::Int4>(f, nf);
	auto starts = rd<int>(f, nv + 1);
	auto idx = rd<int>(f, nnz);
	auto diag = rd<SE::SeMatrix3f>(f, nv);
	auto off = rd<SE::SeMatrix3f>(f, nnz);
	auto ef = rd<SE::EfSet>(f, nRec);
	auto ee = rd<SE::EeSet>(f, nRec);
	auto vf = rd<SE::VfSet>(f, nRec);
	auto r = rd<SE::SeVec3fSimd>(f, nv);
	fclose(f);

	SE::SeCsr<int> nbr(starts, idx);
	// caller-side prefix arrays: only the last element is read (cpp:306-308)
	std::vector<unsigned> efCounts(ne + 1, (unsigned)hdr[4]), eeCounts(ne + 1, (unsigned)hdr[5]), vfCounts(nv + 1, (unsigned)hdr[6]);

	SE::SeSchwarzPreconditioner pre;
	pre.m_positions = pos.data();
	pre.m_edges = edges.data();
	pre.m_faces = faces.data();
	pre.m_neighbours = &nbr;
	pre.AllocatePrecoditioner(nv, ne, nf);
	if (pre.LastStatus() != MAS_OK) return 3;
	pre.PreparePreconditioner(diag.data(), off.data(), starts.data(), ef.data(), ee.data(), vf.data(), efCounts.data(), eeCounts.data(),
		vfCounts.data());
	if (pre.LastStatus() != MAS_OK) return 4;
	std::vector<SE::SeVec3fSimd> z(nv);
	pre.Preconditioning(z.data(), r.data(), 3 * nv);
	if (pre.LastStatus() != MAS_OK) return 5;

	FILE* o = fopen(argv[2], "wb");
	fwrite(z.data(), sizeof(SE::SeVec3fSimd), nv, o);
	fclose(o);
	return 0;
}
