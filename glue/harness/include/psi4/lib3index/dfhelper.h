// HARNESS STAND-IN for psi4/lib3index/dfhelper.h:51-597: the members and methods glue/B200MemDFJK.cc and
// MemDFJK::preiterations touch, with the reference's names and types (tests/test_glue_harness.py compiles one probe
// translation unit against this header AND against the reference's real header, so a drift in either breaks the test).
// initialize() does what DFHelper::initialize does for the in-core STORE method (dfhelper.cc:149-215) from data the
// test injects instead of Libint2: the pair mask stands in for the Schwarz integrals, the packed tensors for
// prepare_AO_core / prepare_AO_wK_core.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "psi4/libmints/basisset.h"
#include "psi4/libmints/matrix.h"
namespace psi {

/// what the test injects per (primary, auxiliary) pair before initialize()
struct HarnessTensors {
    std::vector<unsigned char> keep;          // nbf*nbf pair mask (symmetric, diagonal kept)
    std::vector<double> Ppq, m1Ppq, wPpq;     // packed pQq tensors on the tables of `keep`
    double omega_built = -1.0;                // the omega wPpq was generated with (checked against set_omega)
};
extern HarnessTensors g_harness_tensors;

class DFHelper {
   public:
    DFHelper(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> aux);
    ~DFHelper();
    void set_method(std::string method) { method_ = method; }
    void set_nthreads(size_t nthreads) { nthreads_ = nthreads; }
    void set_memory(size_t doubles) { memory_ = doubles; }
    size_t get_AO_size() { return big_skips_[nbf_]; }
    double ao_sparsity() { return (1.0 - (double)small_skips_[nbf_] / (double)(nbf_ * nbf_)); }
    void set_AO_core(bool core) { AO_core_ = core; }
    bool get_AO_core() { return AO_core_; }
    void set_schwarz_cutoff(double cutoff) { cutoff_ = cutoff; }
    void set_fitting_condition(double condition) { condition_ = condition; }
    void set_wcombine(bool wcombine);  // dfhelper.h:164-169: throws on true
    void set_do_wK(bool do_wK) { do_wK_ = do_wK; }
    void set_omega(double omega) { omega_ = omega; }
    void set_omega_alpha(double alpha) { omega_alpha_ = alpha; }
    void set_omega_beta(double beta) { omega_beta_ = beta; }
    size_t get_naux() { return naux_; }
    void initialize();
    void prepare_sparsity();
    /// the CPU build (dfhelper.cc:3015-3043): never reached behind the glue; the harness throws if it is
    void build_JK(std::vector<SharedMatrix> Cleft, std::vector<SharedMatrix> Cright, std::vector<SharedMatrix> D,
                  std::vector<SharedMatrix> J, std::vector<SharedMatrix> K, std::vector<SharedMatrix> wK, size_t max_nocc,
                  bool do_J, bool do_K, bool do_wK, bool lr_symmetric);

   protected:
    std::shared_ptr<BasisSet> primary_;
    std::shared_ptr<BasisSet> aux_;
    size_t nbf_;
    size_t naux_;
    size_t memory_ = 256000000;
    std::string method_ = "STORE";
    bool AO_core_ = true;
    size_t nthreads_ = 1;
    double cutoff_ = 1e-12;
    double condition_ = 1e-12;
    bool built_ = false;
    bool do_wK_ = false;
    bool wcombine_ = false;
    double omega_ = 0.0;
    double omega_alpha_ = 0.0;
    double omega_beta_ = 0.0;
    bool sparsity_prepared_ = false;
    std::unique_ptr<double[]> Ppq_;
    std::unique_ptr<double[]> wPpq_;
    std::unique_ptr<double[]> m1Ppq_;
    std::vector<size_t> small_skips_;
    std::vector<size_t> big_skips_;
    std::vector<size_t> symm_small_skips_;
    std::vector<size_t> symm_ignored_columns_;
    std::vector<size_t> symm_big_skips_;
    std::vector<size_t> schwarz_fun_index_;
};

}  // namespace psi
