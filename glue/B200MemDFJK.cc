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
    // (a second initialize() on the same object lands here again: set_layout re-initialises the engine, which drops the
    // tensors of the previous pass, and the fresh ones go up -- MemDFJK::preiterations simply recomputes, so do we)
    check(h, b200jk_set_layout(h, nbf_, naux_, small_skips_.data(), big_skips_.data(), schwarz_fun_index_.data()),
          "set_layout");
    layout_sent_ = true;
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
    if (!layout_sent_) {  // after move_to_device the engine already has the tables (and the tensors that hang on them)
        check(h, b200jk_set_layout(h, nbf_, naux_, small_skips_.data(), big_skips_.data(), schwarz_fun_index_.data()),
              "set_layout");
        layout_sent_ = true;
    }
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
    int rc = b200jk_create(&handle_, ngpu_, nullptr);
    if (rc != B200JK_OK) {
        // the handle exists even when creation failed half-way (it carries the message); the destructor will not run
        std::string msg = std::string("B200MemDFJK: create: ") + b200jk_last_error(handle_);
        if (handle_) b200jk_destroy(handle_);
        handle_ = nullptr;
        throw PSIEXCEPTION(msg);
    }
}

std::shared_ptr<B200MemDFJK> B200MemDFJK::build(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> auxiliary,
                                                Options& options, int ngpu, bool release_host) {
    auto jk = std::make_shared<B200MemDFJK>(primary, auxiliary, options, ngpu, release_host);
    jk->set_wcombine(false);  // jk.cc:145-146
    // _set_dfjk_options<MemDFJK>, jk.cc:58-68
    double cutoff = options.get_str("SCREENING") == "NONE" ? 0.0 : options.get_double("INTS_TOLERANCE");
    if (options["INTS_TOLERANCE"].has_changed() || options.get_str("SCREENING") == "NONE") jk->set_cutoff(cutoff);
    if (options["PRINT"].has_changed()) jk->set_print(options.get_int("PRINT"));
    if (options["DEBUG"].has_changed()) jk->set_debug(options.get_int("DEBUG"));
    if (options["BENCH"].has_changed()) jk->set_bench(options.get_int("BENCH"));
    jk->set_condition(options.get_double("DF_FITTING_CONDITION"));
    if (options["DF_INTS_NUM_THREADS"].has_changed()) jk->set_df_ints_num_threads(options.get_int("DF_INTS_NUM_THREADS"));
    if (options["WCOMBINE"].has_changed()) jk->set_wcombine(options.get_bool("WCOMBINE"));  // jk.cc:148
    return jk;
}

B200MemDFJK::~B200MemDFJK() {
    if (handle_) b200jk_destroy(handle_);  // also unregisters the page-locked matrices (pinned_ still holds them)
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
    // D_ao_/J_ao_/K_ao_/wK_ao_ are allocated once by JK::compute_D / allocate_JK / USO2AO (jk.cc:314-446) and reused by
    // every compute(); page-locking them lets the engine DMA in place.  When the matrix count changes psi4 re-creates
    // them, so the set is compared on every build: matrices psi4 no longer uses are unregistered FIRST and only then
    // released (the glue's reference kept them alive, so the memory was never freed while page-locked), new ones
    // are registered.  Anything that cannot be registered is simply staged through the engine's pinned buffers.
    // (sizes from the matrices themselves: basisset.h pulls in <libint2/shell.h>, which this file does not need)
    std::vector<SharedMatrix> now;
    auto collect = [&](std::vector<SharedMatrix>& v) {
        for (auto& m : v)
            if (m) now.push_back(m);
    };
    collect(D_ao_);
    if (do_J_) collect(J_ao_);
    if (do_K_) collect(K_ao_);
    if (do_wK_) collect(wK_ao_);
    if (now == pinned_) return;
    for (auto& old : pinned_) {
        bool keep = false;
        for (auto& n : now) keep = keep || n == old;
        if (!keep && b200jk_unregister_host(handle_, old->get_pointer()) != B200JK_OK) fail("unregister_host");
    }
    for (auto& n : now) {
        bool have = false;
        for (auto& old : pinned_) have = have || n == old;
        const size_t bytes = sizeof(double) * n->rowspi()[0] * n->colspi()[0];
        if (!have && bytes && b200jk_register_host(handle_, n->get_pointer(), bytes) != B200JK_OK) fail("register_host");
    }
    pinned_ = now;  // drops the last reference to the matrices psi4 has already let go of
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
