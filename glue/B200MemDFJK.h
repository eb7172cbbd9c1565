/*
 * psi4-side glue for libb200jk.so: a MemDFJK whose J/K build runs on B200 GPUs.
 *
 * Compiled against a psi4 source/build tree (headers only from psi4: libfock/jk.h, lib3index/dfhelper.h,
 * libmints/matrix.h) and linked with -lb200jk.  Nothing in psi4 is modified: the object is handed to the SCF driver
 * with  psi4.energy('scf', jk=obj)  (psi4/driver/procrouting/proc.py:2043-2045) or selected in the factory with the
 * three-line #ifdef shown in INTEGRATION.md (libfock/jk.cc:143-150).
 *
 * What it replaces: DFHelper::build_JK (lib3index/dfhelper.cc:3015-3438) called from MemDFJK::compute_JK
 * (libfock/MemDFJK.cc:97-111).  What it keeps: everything else of MemDFJK/JK/DFHelper -- options, Schwarz screening,
 * Libint2 integrals, fitting, USO2AO/AO2USO, printing.
 */
#pragma once

#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "psi4/lib3index/dfhelper.h"
#include "psi4/libfock/jk.h"

#include "b200jk.h"

namespace psi {

/// Reads DFHelper's protected tables and in-core tensors (dfhelper.h:333-597) and moves them into HBM.
class B200DFHelper : public DFHelper {
   public:
    B200DFHelper(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> aux) : DFHelper(primary, aux) {}

    /// After DFHelper::initialize(): layout tables + Ppq_ (+ m1Ppq_, wPpq_) -> device, Q-sharded.  With
    /// release_host the host copies are freed (DFHelper::build_JK must then never be called on this object).
    void move_to_device(b200jk_t* h, bool release_host);

    /// HBM bytes per GPU the engine needs for this system (b200jk_hbm_estimate, the device analogue of
    /// get_core_size dfhelper.cc:216-236); requires prepare_sparsity() to have run.
    size_t device_bytes_per_gpu(b200jk_t* h, size_t max_nocc);

    bool tensors_on_host() const { return static_cast<bool>(Ppq_); }

   private:
    bool layout_sent_ = false;  // the engine holds this object's tables (a second set_layout would drop the tensors)
};

class B200MemDFJK : public MemDFJK {
   protected:
    b200jk_t* handle_ = nullptr;
    int ngpu_;
    bool release_host_;
    // page-locked D/J/K/wK matrices.  The glue holds a reference to each while it is registered: psi4 re-creates the
    // matrices whenever the matrix count changes (JK::compute_D / allocate_JK, jk.cc:314-389), and memory must not go
    // back to the allocator while it is still cudaHostRegister'ed.
    std::vector<SharedMatrix> pinned_;

    std::string name() override { return "B200MemDFJK"; }
    void preiterations() override;   // MemDFJK.cc:71-96, then upload
    void compute_JK() override;      // MemDFJK.cc:97-111 on the GPUs
    void postiterations() override;  // frees the device tensors

    void register_persistent_matrices();
    [[noreturn]] void fail(const std::string& where) const;

   public:
    /// ngpu GPUs (device ordinals 0..ngpu-1) driven by this one process; the auxiliary index is sharded over them.
    B200MemDFJK(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> auxiliary, Options& options,
                int ngpu = 1, bool release_host = true);
    ~B200MemDFJK() override;

    /// What JK::build_JK does for SCF_TYPE = MEM_DF (libfock/jk.cc:143-148) with this class in MemDFJK's place:
    /// construct, set_wcombine(false), then _set_dfjk_options (jk.cc:58-68: DF_FITTING_CONDITION, INTS_TOLERANCE /
    /// SCREENING, PRINT, DEBUG, BENCH, DF_INTS_NUM_THREADS) and WCOMBINE.  Both attach points of INTEGRATION.md use
    /// it, so a B200MemDFJK is configured exactly as the stock MemDFJK of the same options would be.
    static std::shared_ptr<B200MemDFJK> build(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> auxiliary,
                                              Options& options, int ngpu = 1, bool release_host = true);

    double condition() const { return condition_; }
    size_t pinned_matrices() const { return pinned_.size(); }

    void print_header() const override;

    /// Per-kernel device timings of the last build (b200jk_get_stats).
    b200jk_stats last_stats() const;
};

}  // namespace psi
