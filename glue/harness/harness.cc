/*
 * harness.cc -- TEST INFRASTRUCTURE: the little of psi4 that glue/B200MemDFJK.cc needs in order to be linked and RUN
 * without a psi4 build (psi4 needs Libint2 / LibXC / gau2grid, none of which exist in this image).
 *
 * It implements the stand-in classes declared under glue/harness/include/psi4/ by restating, for the C1 case, exactly
 * the reference statements the glue sits between:
 *   JK::common_init / compute / compute_D / allocate_JK / USO2AO / zero   libfock/jk.cc:248-277, :595-681, :314-401, :690-703
 *   MemDFJK ctor / preiterations / compute_JK / setters / print_header    libfock/MemDFJK.cc:56-160
 *   DFHelper::prepare_sparsity tables / initialize (in-core STORE)        lib3index/dfhelper.cc:371-416, :149-215
 * and a small extern "C" driver so tests/test_glue_harness.py (ctypes) can play psi4's SCF driver: build the object
 * the way JK::build_JK would, initialize(), push orbitals, compute(), read J()/K()/wK().
 *
 * The Libint2 side is replaced by data the test injects (pair mask = Schwarz screening result, packed tensors =
 * prepare_AO_core output); MemDFJK::compute_JK of the stand-in calls DFHelper::build_JK, which throws: behind the glue
 * the CPU build must never run.
 */
#include <chrono>
#include <cstring>
#include <map>
#include <sstream>

#include "psi4/lib3index/dfhelper.h"
#include "psi4/libfock/jk.h"
#include "psi4/libmints/basisset.h"
#include "psi4/liboptions/liboptions.h"
#include "psi4/libpsi4util/PsiOutStream.h"
#include "psi4/libpsi4util/exception.h"
#include "psi4/libpsi4util/process.h"
#include "psi4/libqt/qt.h"

#include "B200MemDFJK.h"

namespace psi {

std::shared_ptr<PsiOutStream> outfile = std::make_shared<PsiOutStream>();
Process::Environment Process::environment;
HarnessTensors g_harness_tensors;

static std::map<std::string, int> g_timer_calls;
void timer_on(const std::string& key) { g_timer_calls[key]++; }
void timer_off(const std::string&) {}
int harness_timer_count(const std::string& key) { return g_timer_calls.count(key) ? g_timer_calls[key] : 0; }

// ---- DFHelper ---------------------------------------------------------------------------------------------------
DFHelper::DFHelper(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> aux)
    : primary_(primary), aux_(aux), nbf_(primary->nbf()), naux_(aux->nbf()) {}
DFHelper::~DFHelper() {}

void DFHelper::set_wcombine(bool wcombine) {
    if (wcombine) throw PSIEXCEPTION("JK: wcombine option is currently not available.");  // dfhelper.h:164-169
    wcombine_ = wcombine;
}

// the table-building half of prepare_sparsity (dfhelper.cc:371-416); the Schwarz integrals before it are Libint2's and
// are replaced by the injected mask
void DFHelper::prepare_sparsity() {
    if (sparsity_prepared_) return;
    const auto& keep = g_harness_tensors.keep;
    if (keep.size() != nbf_ * nbf_) throw PSIEXCEPTION("harness: no pair mask injected for this basis");
    schwarz_fun_index_.assign(nbf_ * nbf_, 0);
    small_skips_.assign(nbf_ + 1, 0);
    big_skips_.assign(nbf_ + 1, 0);
    symm_small_skips_.assign(nbf_, 0);
    symm_ignored_columns_.assign(nbf_, 0);
    symm_big_skips_.assign(nbf_ + 1, 0);
    size_t coltots = 0;
    for (size_t i = 0; i < nbf_; i++) {
        size_t count = 0;
        for (size_t j = 0; j < nbf_; j++) {
            if (keep[i * nbf_ + j]) {
                count++;
                schwarz_fun_index_[i * nbf_ + j] = count;
            }
        }
        small_skips_[i] = count;
        coltots += count;
    }
    small_skips_[nbf_] = coltots;
    for (size_t j = 0; j < nbf_; j++) big_skips_[j + 1] = big_skips_[j] + naux_ * small_skips_[j];
    for (size_t i = 0; i < nbf_; i++) {
        size_t size = 0, skip = 0;
        for (size_t j = 0; j < nbf_; j++) {
            if (schwarz_fun_index_[i * nbf_ + j]) (j >= i ? size : skip)++;
        }
        symm_small_skips_[i] = size;
        symm_ignored_columns_[i] = skip;
        symm_big_skips_[i + 1] = symm_big_skips_[i] + naux_ * size;
    }
    sparsity_prepared_ = true;
}

// dfhelper.cc:149-215 for method STORE, in core: sparsity, then the packed tensors (prepare_AO_core :514-588 /
// prepare_AO_wK_core :589-699 -- here: the injected data, re-read on every call as the reference recomputes)
void DFHelper::initialize() {
    if (method_ != "STORE") throw PSIEXCEPTION("harness: only method STORE");
    sparsity_prepared_ = false;
    prepare_sparsity();
    AO_core_ = true;
    const size_t n = big_skips_[nbf_];
    auto load = [&](std::unique_ptr<double[]>& dst, const std::vector<double>& src, const char* what) {
        if (src.size() != n) throw PSIEXCEPTION(std::string("harness: injected ") + what + " has the wrong size");
        dst = std::make_unique<double[]>(n);
        std::memcpy(dst.get(), src.data(), n * sizeof(double));
    };
    load(Ppq_, g_harness_tensors.Ppq, "Ppq");
    if (do_wK_) {
        if (g_harness_tensors.omega_built != omega_)
            throw PSIEXCEPTION("harness: wK tensors were generated for a different omega than set_omega() carries");
        load(m1Ppq_, g_harness_tensors.m1Ppq, "m1Ppq");
        load(wPpq_, g_harness_tensors.wPpq, "wPpq");
    }
    built_ = true;
}

void DFHelper::build_JK(std::vector<SharedMatrix>, std::vector<SharedMatrix>, std::vector<SharedMatrix>,
                        std::vector<SharedMatrix>, std::vector<SharedMatrix>, std::vector<SharedMatrix>, size_t, bool, bool,
                        bool, bool) {
    throw PSIEXCEPTION("harness: DFHelper::build_JK (the CPU build) was reached behind the B200 glue");
}

// ---- JK (C1 case) -----------------------------------------------------------------------------------------------
JK::JK(std::shared_ptr<BasisSet> primary) : primary_(primary) { common_init(); }
JK::~JK() {}
void JK::common_init() {  // jk.cc:248-277
    print_ = 1;
    debug_ = 0;
    bench_ = 0;
    memory_ = 32000000L;
    omp_nthread_ = 1;
    cutoff_ = 1.0E-12;
    do_csam_ = false;
    do_J_ = true;
    do_K_ = true;
    do_wK_ = false;
    wcombine_ = false;
    lr_symmetric_ = false;
    omega_ = 0.0;
    omega_alpha_ = 1.0;
    omega_beta_ = 0.0;
}
size_t JK::memory_overhead() const { return 0; }
void JK::set_wcombine(bool wcombine) {  // jk.cc:683-688
    wcombine_ = wcombine;
    if (wcombine) throw PSIEXCEPTION("To combine exchange terms, use MemDFJK\n");
}
void JK::initialize() { preiterations(); }
void JK::finalize() { postiterations(); }

void JK::compute_D() {  // jk.cc:314-354, one irrep
    bool same = C_left_.size() == D_.size();
    if (!same) {
        D_.clear();
        for (size_t N = 0; N < C_left_.size(); ++N) {
            std::stringstream s;
            s << "D " << N << " (SO)";
            D_.push_back(std::make_shared<Matrix>(s.str(), C_left_[N]->rowspi(), C_right_[N]->rowspi(), 0));
        }
    }
    for (size_t N = 0; N < D_.size(); ++N) {
        D_[N]->zero();
        const int nsol = C_left_[N]->rowspi()[0], nocc = C_left_[N]->colspi()[0], nsor = C_right_[N]->rowspi()[0];
        if (!nsol || !nsor || !nocc) continue;
        double* Dp = D_[N]->get_pointer();
        const double* Cl = C_left_[N]->get_pointer();
        const double* Cr = C_right_[N]->get_pointer();
        for (int m = 0; m < nsol; m++)  // C_DGEMM('N','T', nsol, nsor, nocc, ...)
            for (int n = 0; n < nsor; n++) {
                double v = 0.0;
                for (int i = 0; i < nocc; i++) v += Cl[(size_t)m * nocc + i] * Cr[(size_t)n * nocc + i];
                Dp[(size_t)m * nsor + n] = v;
            }
    }
}

void JK::allocate_JK() {  // jk.cc:355-389
    bool same = J_.size() == D_.size();
    if (!same) {
        J_.clear();
        K_.clear();
        wK_.clear();
        auto make = [&](std::vector<SharedMatrix>& v, const char* tag, bool tasked) {
            for (size_t N = 0; N < D_.size() && tasked; ++N) {
                std::stringstream s;
                s << tag << " " << N << " (SO)";
                v.push_back(std::make_shared<Matrix>(s.str(), D_[N]->rowspi(), D_[N]->rowspi(), 0));
            }
        };
        make(J_, "J", do_J_);
        make(K_, "K", do_K_);
        make(wK_, "wK", do_wK_);
    }
}

void JK::USO2AO() {  // jk.cc:390-401: AO2USO_->nirrep() == 1
    allocate_JK();
    C_left_ao_ = C_left_;
    C_right_ao_ = C_right_;
    D_ao_ = D_;
    J_ao_ = J_;
    K_ao_ = K_;
    wK_ao_ = wK_;
}

void JK::zero() {  // jk.cc:690-703
    if (do_J_)
        for (auto& J : J_) J->zero();
    if (do_K_)
        for (auto& K : K_) K->zero();
    if (do_wK_)
        for (auto& wK : wK_) wK->zero();
}

void JK::compute() {  // jk.cc:595-681
    if (C_left_.size() && !C_right_.size()) {
        lr_symmetric_ = true;
        C_right_ = C_left_;
    } else {
        lr_symmetric_ = false;
    }
    if (C_left_.size() != C_right_.size()) throw PSIEXCEPTION("JK: C_left/C_right irrep mismatch!");
    for (size_t i = 0; i < C_left_.size(); i++) {
        if (C_left_[i]->colspi() != C_right_[i]->colspi())
            throw PSIEXCEPTION("JK: C_left/C_right MO zip index size mismatch!");
    }
    timer_on("JK: D");
    compute_D();
    timer_off("JK: D");
    timer_on("JK: USO2AO");
    USO2AO();
    timer_off("JK: USO2AO");
    timer_on("JK: JK");
    compute_JK();
    timer_off("JK: JK");
    if (lr_symmetric_) C_right_.clear();
}

// ---- MemDFJK ----------------------------------------------------------------------------------------------------
MemDFJK::MemDFJK(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> auxiliary, Options& options)
    : JK(primary), options_(options), auxiliary_(auxiliary) {
    common_init();
}
MemDFJK::~MemDFJK() {}
void MemDFJK::common_init() { dfh_ = std::make_shared<DFHelper>(primary_, auxiliary_); }
size_t MemDFJK::memory_estimate() { return 0; }

void MemDFJK::preiterations() {  // MemDFJK.cc:71-96
    dfh_->set_nthreads(omp_nthread_);
    dfh_->set_schwarz_cutoff(cutoff_);
    dfh_->set_method("STORE");
    dfh_->set_fitting_condition(condition_);
    dfh_->set_memory(memory_ - memory_overhead());
    dfh_->set_do_wK(do_wK_);
    dfh_->set_omega(omega_);
    if (do_wK_) {
        dfh_->set_wcombine(wcombine_);
    } else {
        dfh_->set_wcombine(false);
        wcombine_ = false;
    }
    dfh_->set_omega_alpha(omega_alpha_);
    dfh_->set_omega_beta(omega_beta_);
    dfh_->initialize();
}

void MemDFJK::compute_JK() {  // MemDFJK.cc:97-111
    zero();
    dfh_->build_JK(C_left_ao_, C_right_ao_, D_ao_, J_ao_, K_ao_, wK_ao_, max_nocc(), do_J_, do_K_, do_wK_, lr_symmetric_);
    if (lr_symmetric_ && do_wK_)
        for (auto& m : wK_ao_) m->hermitivitize();
}
void MemDFJK::postiterations() {}

void MemDFJK::print_header() const {  // MemDFJK.cc:113-132
    if (print_) {
        outfile->Printf("  ==> MemDFJK: Density-Fitted J/K Matrices <==\n\n");
        outfile->Printf("    J tasked:           %11s\n", (do_J_ ? "Yes" : "No"));
        outfile->Printf("    K tasked:           %11s\n", (do_K_ ? "Yes" : "No"));
        outfile->Printf("    wK tasked:          %11s\n", (do_wK_ ? "Yes" : "No"));
        if (do_wK_) outfile->Printf("    Omega:              %11.3E\n", omega_);
        outfile->Printf("    OpenMP threads:     %11d\n", omp_nthread_);
        outfile->Printf("    Memory [MiB]:       %11ld\n", (memory_ * 8L) / (1024L * 1024L));
        outfile->Printf("    Algorithm:          %11s\n", (dfh_->get_AO_core() ? "Core" : "Disk"));
        outfile->Printf("    Schwarz Cutoff:     %11.0E\n", cutoff_);
        outfile->Printf("    Mask sparsity (%%):  %11.4f\n", 100. * dfh_->ao_sparsity());
        outfile->Printf("    Fitting Condition:  %11.0E\n\n", condition_);
    }
}
int MemDFJK::max_nocc() const {  // MemDFJK.cc:133-139
    int max_nocc = 0;
    for (size_t N = 0; N < C_left_ao_.size(); N++)
        max_nocc = (C_left_ao_[N]->colspi()[0] > max_nocc ? C_left_ao_[N]->colspi()[0] : max_nocc);
    return max_nocc;
}
void MemDFJK::set_omega_alpha(double alpha) {
    omega_alpha_ = alpha;
    dfh_->set_omega_alpha(omega_alpha_);
}
void MemDFJK::set_omega_beta(double beta) {
    omega_beta_ = beta;
    dfh_->set_omega_beta(omega_beta_);
}
void MemDFJK::set_do_wK(bool tf) {
    do_wK_ = tf;
    dfh_->set_do_wK(tf);
}
void MemDFJK::set_wcombine(bool wcombine) {
    wcombine_ = wcombine;
    if (dfh_) dfh_->set_wcombine(wcombine);
}
void MemDFJK::set_cutoff(double cutoff) {
    cutoff_ = cutoff;
    if (dfh_) dfh_->set_schwarz_cutoff(cutoff);
}

}  // namespace psi

// ---- the driver the test plays psi4 with -----------------------------------------------------------------------------
using namespace psi;

namespace {
std::string g_err;
struct Session {
    std::shared_ptr<BasisSet> primary, aux;
    std::shared_ptr<B200MemDFJK> jk;
};
template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

const char* harness_last_error() { return g_err.c_str(); }

/* inject what Libint2 + DFHelper::prepare_AO_core would have produced (packed on the tables of `keep`) */
int harness_inject(int nbf, const unsigned char* keep, size_t packed, const double* Ppq, const double* m1Ppq,
                   const double* wPpq, double omega_built) {
    return guarded([&] {
        auto& t = g_harness_tensors;
        t.keep.assign(keep, keep + (size_t)nbf * nbf);
        t.Ppq.assign(Ppq, Ppq + packed);
        t.m1Ppq.clear();
        t.wPpq.clear();
        if (m1Ppq) t.m1Ppq.assign(m1Ppq, m1Ppq + packed);
        if (wPpq) t.wPpq.assign(wPpq, wPpq + packed);
        t.omega_built = omega_built;
    });
}

/* option keys as psi4's set_options would change them (marks has_changed) */
void harness_set_option_double(const char* key, double v) { Process::environment.options.set_double(key, v); }
void harness_set_option_str(const char* key, const char* v) { Process::environment.options.set_str(key, v); }
void harness_reset_options() { Process::environment.options = Options(); }

/* B200MemDFJK::build == what JK::build_JK does for MEM_DF (jk.cc:143-148) */
void* harness_build(int nbf, int naux, int ngpu, int release_host) {
    Session* s = nullptr;
    int rc = guarded([&] {
        s = new Session();
        s->primary = std::make_shared<BasisSet>(nbf);
        s->aux = std::make_shared<BasisSet>(naux);
        s->jk = B200MemDFJK::build(s->primary, s->aux, Process::environment.options, ngpu, release_host != 0);
    });
    if (rc) {
        delete s;
        return nullptr;
    }
    return s;
}
void harness_destroy(void* p) { delete static_cast<Session*>(p); }

int harness_set_tasks(void* p, int do_J, int do_K, int do_wK, double omega) {
    return guarded([&] {
        auto& jk = static_cast<Session*>(p)->jk;
        jk->set_do_J(do_J != 0);
        jk->set_do_K(do_K != 0);
        jk->set_do_wK(do_wK != 0);
        jk->set_omega(omega);
    });
}
int harness_initialize(void* p) {
    return guarded([&] { static_cast<Session*>(p)->jk->initialize(); });
}
int harness_finalize(void* p) {
    return guarded([&] { static_cast<Session*>(p)->jk->finalize(); });
}
double harness_condition(void* p) { return static_cast<Session*>(p)->jk->condition(); }
double harness_cutoff(void* p) { return static_cast<Session*>(p)->jk->get_cutoff(); }
int harness_pinned(void* p) { return (int)static_cast<Session*>(p)->jk->pinned_matrices(); }
int harness_tensors_on_host(void* p) {
    return std::static_pointer_cast<B200DFHelper>(static_cast<Session*>(p)->jk->dfh())->tensors_on_host() ? 1 : 0;
}
int harness_timer_calls(const char* key) { return harness_timer_count(key); }

/* one SCF-iteration's worth of driver code: C_left().clear(); push_back; (C_right likewise or left empty); compute();
 * then the J()/K()/wK() the driver would re-fetch (jk.h:149-159), copied out. */
int harness_compute(void* p, int nmat, int nbf, const int* nocc, const double* const* Cl, const double* const* Cr,
                    double* const* J, double* const* K, double* const* wK) {
    return guarded([&] {
        auto& jk = static_cast<Session*>(p)->jk;
        jk->C_left().clear();
        jk->C_right().clear();
        for (int i = 0; i < nmat; i++) {
            auto c = std::make_shared<Matrix>("C_left", nbf, nocc[i]);
            std::memcpy(c->get_pointer(), Cl[i], sizeof(double) * nbf * nocc[i]);
            jk->C_left().push_back(c);
            if (Cr) {
                auto r = std::make_shared<Matrix>("C_right", nbf, nocc[i]);
                std::memcpy(r->get_pointer(), Cr[i], sizeof(double) * nbf * nocc[i]);
                jk->C_right().push_back(r);
            }
        }
        jk->compute();
        const size_t n2 = sizeof(double) * nbf * nbf;
        for (int i = 0; i < nmat; i++) {
            if (J && i < (int)jk->J().size()) std::memcpy(J[i], jk->J()[i]->get_pointer(), n2);
            if (K && i < (int)jk->K().size()) std::memcpy(K[i], jk->K()[i]->get_pointer(), n2);
            if (wK && i < (int)jk->wK().size()) std::memcpy(wK[i], jk->wK()[i]->get_pointer(), n2);
        }
    });
}

/* print_header() text (MemDFJK's block followed by the engine's) */
int harness_header(void* p, char* out, size_t cap) {
    return guarded([&] {
        outfile->clear();
        static_cast<Session*>(p)->jk->print_header();
        std::strncpy(out, outfile->text().c_str(), cap - 1);
        out[cap - 1] = 0;
    });
}
int harness_last_launches(void* p) { return (int)static_cast<Session*>(p)->jk->last_stats().launches; }
}
