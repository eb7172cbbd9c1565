/*
 * psi4-side glue for libb200jk.so (see B200MemDFJK.h).  Reference statements it stands in for are cited per function.
 */
#include "B200MemDFJK.h"

#include <cstdint>

#include "psi4/libmints/matrix.h"
#include "psi4/liboptions/liboptions.h"
#include "psi4/libpsi4util/PsiOutStream.h"
#include "psi4/libpsi4util/exception.h"
#include "psi4/libqt/qt.h"

namespace psi {

namespace {
void check(b200jk_t* h, int rc, const char* where) {
    if (rc == B200JK_OK) return;
    std::string msg = std::string("B200MemDFJK: ") + where + ": " + b200jk_last_error(h);
    throw PSIEXCEPTION(msg);  // the engine never falls back to the CPU path; neither does the glue
}
}  // namespace

// ---- B200DFHelper ---------------------------------------------------------------------------------------------

void B200DFHelper::move_to_device(b200jk_t* h, bool release_host) {
    if (!AO_core_) {
        // the reference would switch to the disk sub-algorithm here (dfhelper.cc:246-249); on the GPUs the
        // analogue is more Q shards, so refuse instead of degrading (SCF_SUBTYPE=INCORE semantics, :259-262)
        throw PSIEXCEPTION("B200MemDFJK: DFHelper chose the out-of-core algorithm; set SCF_SUBTYPE INCORE / raise memory");
    }
    check(h, b200jk_set_layout(h, nbf_, naux_, small_skips_.data(), big_skips_.data(), schwarz_fun_index_.data()),
          "set_layout");
    // pageable memory is fine here: the engine copies it through its page-locked ring with a few threads while the
    // DMA engines drain it (measured 16-32 GB/s on the B200 box; page-locking the whole tensor first costs more than
    // it saves for a one-shot upload)
    check(h, b200jk_upload(h, B200JK_TENSOR_PPQ, Ppq_.get()), "upload Ppq");
    if (do_wK_) {
        // dfhelper.cc:589-699: m1Ppq_ = J^-1 (A|mn), wPpq_ = (A|erf(w r)/r|mn), same pQq layout
        check(h, b200jk_upload(h, B200JK_TENSOR_M1PPQ, m1Ppq_.get()), "upload m1Ppq");
        check(h, b200jk_upload(h, B200JK_TENSOR_WPPQ, wPpq_.get()), "upload wPpq");
    }
    if (release_host) {
        Ppq_.reset();
        m1Ppq_.reset();
        wPpq_.reset();
    }
}

size_t B200DFHelper::device_bytes_per_gpu(b200jk_t* h, size_t max_nocc) {
    if (!sparsity_prepared_) prepare_sparsity();
    check(h, b200jk_set_layout(h, nbf_, naux_, small_skips_.data(), big_skips_.data(), schwarz_fun_index_.data()),
          "set_layout");
    uint64_t bytes = 0;
    check(h, b200jk_hbm_estimate(h, max_nocc, do_wK_ ? 1 : 0, &bytes), "hbm_estimate");
    return static_cast<size_t>(bytes);
}

// ---- B200MemDFJK ----------------------------------------------------------------------------------------------

B200MemDFJK::B200MemDFJK(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> auxiliary, Options& options,
                         int ngpu, bool release_host)
    : MemDFJK(primary, auxiliary, options), ngpu_(ngpu), release_host_(release_host) {
    // common_init (MemDFJK.cc:64) made a plain DFHelper; swap in the subclass that can reach the tables
    dfh_ = std::make_shared<B200DFHelper>(primary, auxiliary);
    check(nullptr, b200jk_create(&handle_, ngpu_, nullptr), "create");
}

B200MemDFJK::~B200MemDFJK() {
    if (handle_) b200jk_destroy(handle_);  // also unregisters any page-locked matrices
}

void B200MemDFJK::fail(const std::string& where) const {
    throw PSIEXCEPTION("B200MemDFJK: " + where + ": " + b200jk_last_error(handle_));
}

void B200MemDFJK::preiterations() {
    MemDFJK::preiterations();  // knobs -> DFHelper, sparsity, metric, Libint2 (A|mn), fitting (MemDFJK.cc:71-96)
    timer_on("JK: B200 upload");
    std::static_pointer_cast<B200DFHelper>(dfh_)->move_to_device(handle_, release_host_);
    timer_off("JK: B200 upload");
}

void B200MemDFJK::register_persistent_matrices() {
    // D_ao_/J_ao_/K_ao_/wK_ao_ are allocated once by JK::allocate_JK / USO2AO (jk.cc:355-446) and reused by every
    // compute(); page-locking them lets the engine DMA in place.  When the matrix count changes psi4 re-allocates
    // them, so the set of pointers is compared on every build: vanished ones are unregistered, new ones registered
    // (anything left unregistered is simply staged through the engine's pinned buffers).
    // (sizes from the matrices themselves: basisset.h pulls in <libint2/shell.h>, which this file does not need)
    std::vector<std::pair<double*, size_t>> now;
    auto collect = [&](std::vector<SharedMatrix>& v) {
        for (auto& m : v)
            if (m) now.push_back({m->get_pointer(), sizeof(double) * m->rowspi()[0] * m->colspi()[0]});
    };
    collect(D_ao_);
    if (do_J_) collect(J_ao_);
    if (do_K_) collect(K_ao_);
    if (do_wK_) collect(wK_ao_);
    if (now == pinned_) return;
    for (auto& old : pinned_) {
        bool keep = false;
        for (auto& n : now) keep = keep || n == old;
        if (!keep) b200jk_unregister_host(handle_, old.first);  // already freed by psi4: best effort
    }
    for (auto& n : now)
        if (b200jk_register_host(handle_, n.first, n.second) != B200JK_OK) fail("register_host");
    pinned_ = now;
}

void B200MemDFJK::compute_JK() {
    const int nmat = static_cast<int>(C_left_ao_.size());
    register_persistent_matrices();

    std::vector<const double*> Cl(nmat), Cr(nmat), D(nmat);
    std::vector<double*> J(nmat, nullptr), K(nmat, nullptr), wK(nmat, nullptr);
    std::vector<int> nocc(nmat);
    for (int i = 0; i < nmat; i++) {
        Cl[i] = C_left_ao_[i]->get_pointer();
        Cr[i] = C_right_ao_[i]->get_pointer();
        D[i] = D_ao_[i]->get_pointer();
        nocc[i] = C_left_ao_[i]->colspi()[0];  // dfhelper.cc:3354
        if (do_J_) J[i] = J_ao_[i]->get_pointer();
        if (do_K_) K[i] = K_ao_[i]->get_pointer();
        if (do_wK_) wK[i] = wK_ao_[i]->get_pointer();
    }
    // zero() + build_JK + hermitivitize(wK) of MemDFJK.cc:97-111 all happen inside the engine (outputs overwritten)
    timer_on("JK: B200 build");
    int rc = b200jk_compute(handle_, nmat, Cl.data(), lr_symmetric_ ? nullptr : Cr.data(), nocc.data(), D.data(),
                            J.data(), K.data(), wK.data(), do_J_ ? 1 : 0, do_K_ ? 1 : 0, do_wK_ ? 1 : 0);
    timer_off("JK: B200 build");
    if (rc != B200JK_OK) fail("compute");
}

void B200MemDFJK::postiterations() {
    // MemDFJK::postiterations is empty (MemDFJK.cc:112); the HBM tensors go with the handle in the destructor so a
    // JK object reused across energy() calls (scf_iterator.py:153-155) keeps them.
}

void B200MemDFJK::print_header() const {
    MemDFJK::print_header();  // MemDFJK.cc:113-132
    if (print_) {
        b200jk_stats st{};
        b200jk_get_stats(handle_, &st);
        outfile->Printf("  ==> B200 DF-JK engine <==\n\n");
        outfile->Printf("    GPUs (Q shards):    %11d\n", ngpu_);
        outfile->Printf("    HBM tensors [GiB]:  %11.3f\n", st.hbm_tensor_bytes / 1073741824.0);
        outfile->Printf("    HBM work    [GiB]:  %11.3f\n\n", st.hbm_work_bytes / 1073741824.0);
    }
}

b200jk_stats B200MemDFJK::last_stats() const {
    b200jk_stats st{};
    if (b200jk_get_stats(handle_, &st) != B200JK_OK) fail("get_stats");
    return st;
}

}  // namespace psi
