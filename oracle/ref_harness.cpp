// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// C-ABI harness around the UNMODIFIED reference class SE::SeSchwarzPreconditioner
// (SeSchwarzPreconditioner.h:37-178), compiled by oracle/build_ref.sh into
// oracle/_ref/libmas_ref.so.  It drives the three public methods
// (h:56, h:59-60, h:63) and exposes the private state (h:67-115) read-only so
// the parity tests can compare integer structures bit-for-bit and dense
// inverses by tolerance.  `#define private public` is the only liberty taken.
#include <cstdint>
#include <cstring>
#include <vector>

#define private public
#include "SeSchwarzPreconditioner.h"
#undef private

using namespace SE;

extern int CPU_THREAD_NUM;  // SeOmp.cpp:29-33

namespace {
struct RefHandle
{
	SeSchwarzPreconditioner pre;
	SeCsr<int>* csr = nullptr;
	std::vector<SeVec3fSimd> positions;
	std::vector<Int4> edges, faces;
	std::vector<unsigned int> efCounts, eeCounts, vfCounts;
	int nv = 0, ne = 0, nf = 0;
};
constexpr int kTri = (1 + 96) * 96 / 2 + 16 * 3;  // cpp:165
}

extern "C" {

int ref_sizeof(int what)
{
	switch (what)
	{
	case 0: return (int)sizeof(SeVec3fSimd);
	case 1: return (int)sizeof(SeMatrix3f);
	case 2: return (int)sizeof(Int4);
	case 3: return (int)sizeof(EfSet);
	case 4: return (int)sizeof(EeSet);
	case 5: return (int)sizeof(VfSet);
	case 6: return (int)sizeof(Stencil);
	case 7: return (int)sizeof(Int5);
	case 8: return (int)sizeof(SeMorton64);
	}
	return -1;
}

void ref_set_threads(int n) { CPU_THREAD_NUM = n > 0 ? n : 1; }
int ref_get_threads() { return CPU_THREAD_NUM; }

unsigned long long ref_morton_encode(float x, float y, float z)
{
	SeMorton64 m;
	m.Encode(x, y, z);  // SeMorton.h:75-86
	return (unsigned long long)m;
}

void* ref_create() { return new RefHandle(); }

void ref_destroy(void* h)
{
	auto* r = (RefHandle*)h;
	delete r->csr;
	delete r;
}

// positions: nv*4 floats (xyzw); edges: ne*4 ints; faces: nf*4 ints;
// nbrStarts[nv+1], nbrIdx[nnz] (vertex adjacency CSR without self).
void ref_allocate(void* h, int nv, int ne, int nf, const float* positions, const int* edges, const int* faces,
	const int* nbrStarts, const int* nbrIdx)
{
	auto* r = (RefHandle*)h;
	r->nv = nv; r->ne = ne; r->nf = nf;
	r->positions.resize(nv);
	std::memcpy((void*)r->positions.data(), positions, sizeof(float) * 4 * (size_t)nv);
	r->edges.resize(ne > 0 ? ne : 1);
	r->faces.resize(nf > 0 ? nf : 1);
	if (ne > 0) std::memcpy((void*)r->edges.data(), edges, sizeof(int) * 4 * (size_t)ne);
	if (nf > 0) std::memcpy((void*)r->faces.data(), faces, sizeof(int) * 4 * (size_t)nf);
	std::vector<int> starts(nbrStarts, nbrStarts + nv + 1);
	std::vector<int> idxs(nbrIdx, nbrIdx + starts.back());
	std::vector<int> values;
	delete r->csr;
	r->csr = new SeCsr<int>(starts, idxs, values);
	r->pre.m_positions = r->positions.data();
	r->pre.m_edges = r->edges.data();
	r->pre.m_faces = r->faces.data();
	r->pre.m_neighbours = r->csr;
	r->pre.AllocatePrecoditioner(nv, ne, nf);
}

// diag: nv*9 floats, offdiag: nnz*9 floats, ranges: nv+1 ints (== nbrStarts).
// ef/ee/vf: raw 48-byte records laid out exactly as the reference reads them
// (global stencil index, cpp:328/357/383); totals are what cpp:306-308 read
// from the last element of the caller's prefix arrays.
void ref_prepare(void* h, const float* diag, const float* offdiag, const int* ranges,
	const void* ef, const void* ee, const void* vf, unsigned efTotal, unsigned eeTotal, unsigned vfTotal)
{
	auto* r = (RefHandle*)h;
	r->efCounts.assign((size_t)r->ne + 1, 0u); r->efCounts[r->ne] = efTotal;
	r->eeCounts.assign((size_t)r->ne + 1, 0u); r->eeCounts[r->ne] = eeTotal;
	r->vfCounts.assign((size_t)r->nv + 1, 0u); r->vfCounts[r->nv] = vfTotal;
	r->pre.PreparePreconditioner((const SeMatrix3f*)diag, (const SeMatrix3f*)offdiag, ranges,
		(const EfSet*)ef, (const EeSet*)ee, (const VfSet*)vf,
		r->efCounts.data(), r->eeCounts.data(), r->vfCounts.data());
}

void ref_apply(void* h, float* z, const float* residual)
{
	auto* r = (RefHandle*)h;
	r->pre.Preconditioning((SeVec3fSimd*)z, (const SeVec3fSimd*)residual, 3 * r->nv);
}

// ---- introspection ------------------------------------------------------

int ref_num_level(void* h) { return ((RefHandle*)h)->pre.m_numLevel; }
int ref_total_sz(void* h) { return ((RefHandle*)h)->pre.m_totalSz; }
int ref_total_clusters(void* h) { return ((RefHandle*)h)->pre.m_totalNumberClusters; }
int ref_stencil_num(void* h) { return ((RefHandle*)h)->pre.m_stencilNum; }
int ref_max_neighbours(void* h) { return (int)((RefHandle*)h)->pre.m_mappedNeighbors.Rows(); }

void ref_get_aabb(void* h, float* lower4, float* upper4)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(lower4, &p.m_aabb.Lower, 16);
	std::memcpy(upper4, &p.m_aabb.Upper, 16);
}

void ref_get_level_size(void* h, int* out /* (numLevel+1)*2 */)
{
	auto& p = ((RefHandle*)h)->pre;
	for (int l = 0; l <= p.m_numLevel; ++l) { out[2 * l] = p.m_levelSize[l].x; out[2 * l + 1] = p.m_levelSize[l].y; }
}

void ref_get_morton(void* h, unsigned long long* out)
{
	auto& p = ((RefHandle*)h)->pre;
	for (int i = 0; i < p.m_numVerts; ++i) out[i] = (unsigned long long)p.m_mortonCode[i];
}

void ref_get_sorted_get_original(void* h, int* out)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, p.m_MapperSortedGetOriginal.data(), sizeof(int) * (size_t)p.m_numVerts);
}

void ref_get_original_get_sorted(void* h, int* out)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, p.m_mapperOriginalGetSorted.data(), sizeof(int) * (size_t)p.m_numVerts);
}

// goingNext[0..count): count <= numLevel*nv (h:97, cpp:156)
void ref_get_going_next(void* h, int* out, int count)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, p.m_goingNext.data(), sizeof(int) * (size_t)count);
}

void ref_get_coarse_tables(void* h, int* out /* nv*4 */)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, p.m_coarseTables.data(), sizeof(int) * 4 * (size_t)p.m_numVerts);
}

void ref_get_coarse_space_table(void* h, int level, int* out /* nv */)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, p.m_CoarseSpaceTables[level], sizeof(int) * (size_t)p.m_numVerts);
}

void ref_get_fine_connect_mask(void* h, unsigned* out)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, p.m_fineConnectMask.data(), sizeof(unsigned) * (size_t)p.m_numVerts);
}

void ref_get_mapped_neighbors(void* h, int* num /* nv */, int* table /* rows*nv */)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(num, p.m_mappedNeighborsNum.data(), sizeof(int) * (size_t)p.m_numVerts);
	for (size_t k = 0; k < p.m_mappedNeighbors.Rows(); ++k)
		std::memcpy(table + k * (size_t)p.m_numVerts, p.m_mappedNeighbors[k], sizeof(int) * (size_t)p.m_numVerts);
}

// stencils as 80-byte records + mapped indices (5 ints each)
void ref_get_stencils(void* h, void* stencils, int* mapped)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(stencils, (const void*)p.m_stencils.data(), sizeof(Stencil) * (size_t)p.m_stencilNum);
	std::memcpy(mapped, (const void*)p.m_stencilIndexMapped.data(), sizeof(Int5) * (size_t)p.m_stencilNum);
}

// dense 96x96 Hessian of 32-node block b, row-major, as LDLtInverse512 gathers it
// (cpp:1357-1377) but WITHOUT the identity substitution for padding nodes.
void ref_get_dense_hessian(void* h, int block, float* out)
{
	auto& p = ((RefHandle*)h)->pre;
	for (int x = 0; x < 32; ++x)
		for (int y = 0; y < 32; ++y)
		{
			const SeMatrix3f& t = p.m_hessian32[y][x + block * 32];
			for (int i = 0; i < 3; ++i)
				for (int j = 0; j < 3; ++j)
					out[(x * 3 + i) * 96 + (y * 3 + j)] = t(i, j);
		}
}

// dense symmetric 96x96 inverse of block b, unpacked from m_invSymR following
// the write order of cpp:1435-1495.
void ref_get_dense_inverse(void* h, int block, float* out)
{
	auto& p = ((RefHandle*)h)->pre;
	const float* src = p.m_invSymR.data() + (size_t)block * kTri;
	std::memset(out, 0, sizeof(float) * 96 * 96);
	int off = 0;
	for (int i = 0; i < 96; ++i) out[i * 96 + i] = src[off++];
	for (int it = 0; it < 12; ++it)
	{
		int xBg = it * 8;
		for (int scan = xBg + 1; scan <= 96 - 8; ++scan)
			for (int l = 0; l < 8; ++l)
			{
				float v = src[off++];
				out[(scan + l) * 96 + (xBg + l)] = v;
				out[(xBg + l) * 96 + (scan + l)] = v;
			}
	}
	for (int it = 0; it < 12; ++it)
		for (int lane = 0; lane < 7; ++lane)
		{
			int xBg = it * 8 + lane;
			for (int hh = 96 - 7 + lane; hh < 96; ++hh)
			{
				float v = src[off++];
				out[hh * 96 + xBg] = v;
				out[xBg * 96 + hh] = v;
			}
		}
}

void ref_get_mapped_r(void* h, float* out, int count)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, (const void*)p.m_mappedR.data(), 16 * (size_t)count);
}

void ref_get_mapped_z(void* h, float* out, int count)
{
	auto& p = ((RefHandle*)h)->pre;
	std::memcpy(out, (const void*)p.m_mappedZ.data(), 16 * (size_t)count);
}

}  // extern "C"
