// SeSchwarzPreconditioner.h — C++ drop-in for SE::SeSchwarzPreconditioner over the C ABI of libmas_b200.so.
//
// Same class name, namespace, public members and method signatures as the reference declaration
// (reference SeSchwarzPreconditioner.h:37-63), so a caller's PCG loop recompiles unchanged and links against
// libmas_b200.so instead of SeSchwarzPreconditioner.cpp.  Everything private in the reference (its ~35 std::vector
// buffers and 29 member functions, h:67-178) lives on the GPU behind the opaque mas_handle_t.
//
// Two ways to get the boundary types:
//   * inside the reference's source tree: compile with -DMAS_USE_REFERENCE_HEADERS; the reference's own
//     SeVectorSimd.h / SeMatrix.h / SeCsr.h / SeCollisionElements.h are included and used as they are;
//   * stand-alone (default): layout-identical PODs with the same names are declared below
//     (sizes/offsets measured through the reference headers; tests/test_oracle_vs_reference.py
//     checks them against the compiled reference).
//
// Differences a caller can observe:
//   * results: FP32, within the tolerance stated in DESIGN.md; integer structures bit-exact;
//   * errors: the reference's methods return void and cannot fail; here a failure (no sm_100 GPU, CUDA error,
//     wrong call order) is printed to stderr and kept in LastStatus()/LastError() — there is NO CPU fallback;
//   * extras (not in the reference): device-pointer entry points, stream selection, explicit status.
#ifndef MAS_SE_SCHWARZ_PRECONDITIONER_H
#define MAS_SE_SCHWARZ_PRECONDITIONER_H

#include <cstdio>

#include "mas_b200.h"

#ifdef MAS_USE_REFERENCE_HEADERS
#include "SeCsr.h"
#include "SeVectorSimd.h"
#include "SeMatrix.h"
#include "SeCollisionElements.h"
#else
#include <vector>
namespace SE {

struct alignas(16) SeVec3fSimd { float x = 0.f, y = 0.f, z = 0.f, w = 0.f; };   // 16 B / align 16 (SeVectorSimd.h:45-103)
struct alignas(16) Int4 { int x = 0, y = 0, z = 0, w = 0; };                  // 16 B / align 16 (SeVector.h)
struct Float2Pod { float x, y; };
struct Float3Pod { float x, y, z; };
struct SeMatrix3f { float m_data[9]; };   // 36 B, column-major: (i,j) -> m_data[3*j+i] (SeMatrix.h:681-682)

// 48-byte collision records (SeCollisionElements.h:33-58); field offsets in the comments
struct alignas(16) EfSet { int m_eId; int m_fId; float stiff; Float3Pod m_bary; float pad_[2]; SeVec3fSimd m_normal; };             // 0 4 8 12 | 32
struct alignas(16) EeSet { int m_eId0; int m_eId1; float stiff; float pad0_; Float2Pod m_bary; float pad1_[2]; SeVec3fSimd m_normal; };  // 0 4 8 | 16 | 32
struct alignas(16) VfSet { int m_vId; int m_fId; float stiff; float pad0_; Float2Pod m_bary; float pad1_[2]; SeVec3fSimd m_normal; };    // 0 4 8 | 16 (|24 read by cpp:399) | 32
static_assert(sizeof(SeVec3fSimd) == 16 && sizeof(Int4) == 16 && sizeof(SeMatrix3f) == 36, "boundary layouts");
static_assert(sizeof(EfSet) == 48 && sizeof(EeSet) == 48 && sizeof(VfSet) == 48, "collision record layouts");

// minimal adjacency container with the two accessors this path uses (SeCsr.h:129-142)
template <typename T>
class SeCsr
{
public:
	SeCsr() = default;
	SeCsr(std::vector<int> starts, std::vector<int> idxs) : m_starts(std::move(starts)), m_idxs(std::move(idxs)) {}
	int Size() const { return m_starts.empty() ? 0 : m_starts.back(); }
	int Size(int id) const { return m_starts[id + 1] - m_starts[id]; }
	const int* StartPtr(int id) const { return m_starts.data() + id; }
	const int* IdxPtr(int id) const { return m_idxs.data() + m_starts[id]; }
protected:
	std::vector<int> m_starts, m_idxs;
	std::vector<T> m_values;
};

}  // namespace SE
#endif  // MAS_USE_REFERENCE_HEADERS

namespace SE {

class SeSchwarzPreconditioner
{
public:
	//==== input data (same names as the reference, h:44-51); host pointers owned by the caller
	const SeVec3fSimd* m_positions = nullptr;
	const Int4* m_edges = nullptr;
	const Int4* m_faces = nullptr;
	const SeCsr<int>* m_neighbours = nullptr;

	explicit SeSchwarzPreconditioner(int cudaDevice = 0) : m_device(cudaDevice) {}
	~SeSchwarzPreconditioner() { if (m_handle) mas_destroy(m_handle); }
	SeSchwarzPreconditioner(const SeSchwarzPreconditioner&) = delete;
	SeSchwarzPreconditioner& operator=(const SeSchwarzPreconditioner&) = delete;

	//==== call before time integration once a frame (h:56; the sort runs once per object like cpp:44-64)
	void AllocatePrecoditioner(int numVerts, int numEdges, int numFaces)
	{
		if (!Ensure()) return;
		Note(mas_allocate(m_handle, numVerts, numEdges, numFaces, reinterpret_cast<const float*>(m_positions),
			reinterpret_cast<const int*>(m_edges), reinterpret_cast<const int*>(m_faces),
			m_neighbours ? m_neighbours->StartPtr(0) : nullptr, m_neighbours ? m_neighbours->IdxPtr(0) : nullptr, MAS_MEM_HOST),
			"AllocatePrecoditioner");
		m_numVerts = numVerts; m_numEdges = numEdges;
	}

	//==== call before PCG iteration loop (h:59-60).  The counts arrays are read exactly where cpp:306-308 reads them.
	void PreparePreconditioner(const SeMatrix3f* diagonal, const SeMatrix3f* csrOffDiagonals, const int* csrRanges,
		const EfSet* efSets, const EeSet* eeSets, const VfSet* vfSets, unsigned int* efCounts, unsigned int* eeCounts, unsigned int* vfCounts)
	{
		if (!Ensure()) return;
		const unsigned ef = efCounts ? efCounts[m_numEdges] : 0u, ee = eeCounts ? eeCounts[m_numEdges] : 0u, vf = vfCounts ? vfCounts[m_numVerts] : 0u;
		Note(mas_prepare(m_handle, reinterpret_cast<const float*>(diagonal), reinterpret_cast<const float*>(csrOffDiagonals), csrRanges,
			efSets, eeSets, vfSets, ef, ee, vf, MAS_MEM_HOST), "PreparePreconditioner");
	}

	//==== call during PCG iterations (h:63); dim is ignored exactly as in cpp:100-110
	void Preconditioning(SeVec3fSimd* z, const SeVec3fSimd* residual, int /*dim*/)
	{
		if (!Ensure()) return;
		Note(mas_apply(m_handle, reinterpret_cast<float*>(z), reinterpret_cast<const float*>(residual), MAS_MEM_HOST), "Preconditioning");
	}

	//==== extensions (not in the reference): keep r and z resident in HBM, choose the stream, read the status
	void PreparePreconditionerDevice(const float* diagonal, const float* csrOffDiagonals, const int* csrRanges,
		const void* efSets, const void* eeSets, const void* vfSets, unsigned efTotal, unsigned eeTotal, unsigned vfTotal)
	{
		if (!Ensure()) return;
		Note(mas_prepare(m_handle, diagonal, csrOffDiagonals, csrRanges, efSets, eeSets, vfSets, efTotal, eeTotal, vfTotal, MAS_MEM_DEVICE),
			"PreparePreconditionerDevice");
	}
	void PreconditioningDevice(float* zDevice, const float* residualDevice)
	{
		if (!Ensure()) return;
		Note(mas_apply(m_handle, zDevice, residualDevice, MAS_MEM_DEVICE), "PreconditioningDevice");
	}
	void SetStream(void* cudaStream) { if (Ensure()) Note(mas_set_stream(m_handle, cudaStream), "SetStream"); }
	// waits for the stream; on a sharded context also reports a peer exchange that timed out (see mas_synchronize)
	void Synchronize() { if (Ensure()) Note(mas_synchronize(m_handle), "Synchronize"); }
	void SetOption(int key, int value) { if (Ensure()) Note(mas_set_option(m_handle, key, value), "SetOption"); }
	int LastStatus() const { return m_status; }
	const char* LastError() const { return m_handle ? mas_last_error(m_handle) : "no usable sm_100 GPU (there is no CPU fallback)"; }
	mas_handle_t Handle() { Ensure(); return m_handle; }

private:
	bool Ensure()
	{
		if (m_handle) return true;
		m_status = mas_create(&m_handle, m_device);
		if (m_status != MAS_OK)
		{
			m_handle = nullptr;
			std::fprintf(stderr, "SeSchwarzPreconditioner: mas_create(device %d) failed (%d): %s\n", m_device, m_status, LastError());
			return false;
		}
		return true;
	}
	void Note(int rc, const char* what)
	{
		m_status = rc;
		if (rc != MAS_OK) std::fprintf(stderr, "SeSchwarzPreconditioner::%s failed (%d): %s\n", what, rc, LastError());
	}

	mas_handle_t m_handle = nullptr;
	int m_device = 0;
	int m_status = MAS_OK;
	int m_numVerts = 0, m_numEdges = 0;
};

}  // namespace SE

#endif  // MAS_SE_SCHWARZ_PRECONDITIONER_H
