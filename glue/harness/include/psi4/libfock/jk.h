// HARNESS STAND-IN for psi4/libfock/jk.h:232-589 (class JK) and :1124-1212 (class MemDFJK): same member names, types,
// virtuals and access as the reference (probed against the real header by tests/test_glue_harness.py), with the C1
// branch of JK::compute / compute_D / allocate_JK / USO2AO (jk.cc:314-401, :595-681) and MemDFJK's own methods
// (MemDFJK.cc:56-160) implemented in glue/harness/harness.cc.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "psi4/libmints/matrix.h"
namespace psi {
class BasisSet;
class DFHelper;
class Options;

class JK {
   protected:
    int print_;
    int debug_;
    int bench_;
    size_t memory_;
    int omp_nthread_;
    double cutoff_;
    double do_csam_;
    bool do_J_;
    bool do_K_;
    bool do_wK_;
    bool wcombine_;
    double omega_;
    double omega_alpha_;
    double omega_beta_;
    bool lr_symmetric_;
    std::vector<SharedMatrix> C_left_;
    std::vector<SharedMatrix> C_right_;
    std::vector<SharedMatrix> D_;
    std::vector<SharedMatrix> J_;
    std::vector<SharedMatrix> K_;
    std::vector<SharedMatrix> wK_;
    std::shared_ptr<BasisSet> primary_;
    std::vector<SharedMatrix> C_left_ao_;
    std::vector<SharedMatrix> C_right_ao_;
    std::vector<SharedMatrix> D_ao_;
    std::vector<SharedMatrix> J_ao_;
    std::vector<SharedMatrix> K_ao_;
    std::vector<SharedMatrix> wK_ao_;

    virtual void preiterations() = 0;
    virtual void compute_JK() = 0;
    virtual void postiterations() = 0;
    void common_init();
    size_t memory_overhead() const;
    void compute_D();
    void USO2AO();
    void allocate_JK();
    void zero();

   public:
    JK(std::shared_ptr<BasisSet> primary);
    virtual ~JK();
    virtual bool C1() const = 0;
    virtual std::string name() = 0;
    virtual size_t memory_estimate() = 0;
    virtual void set_cutoff(double cutoff) { cutoff_ = cutoff; }
    double get_cutoff() const { return cutoff_; }
    void set_memory(size_t memory) { memory_ = memory; }
    void set_omp_nthread(int omp_nthread) { omp_nthread_ = omp_nthread; }
    void set_print(int print) { print_ = print; }
    void set_debug(int debug) { debug_ = debug; }
    void set_bench(int bench) { bench_ = bench; }
    void set_do_J(bool do_J) { do_J_ = do_J; }
    virtual void set_do_K(bool do_K) { do_K_ = do_K; }
    virtual void set_do_wK(bool do_wK) { do_wK_ = do_wK; }
    bool get_do_wK() { return do_wK_; }
    virtual void set_wcombine(bool wcombine);
    void set_omega(double omega) { omega_ = omega; }
    double get_omega() { return omega_; }
    virtual void set_omega_alpha(double alpha) { omega_alpha_ = alpha; }
    virtual void set_omega_beta(double beta) { omega_beta_ = beta; }
    void initialize();
    void compute();
    void finalize();
    std::shared_ptr<BasisSet> basisset() { return primary_; }
    std::vector<SharedMatrix>& C_left() { return C_left_; }
    std::vector<SharedMatrix>& C_right() { return C_right_; }
    const std::vector<SharedMatrix>& J() const { return J_; }
    const std::vector<SharedMatrix>& K() const { return K_; }
    const std::vector<SharedMatrix>& wK() const { return wK_; }
    const std::vector<SharedMatrix>& D() const { return D_; }
    virtual void print_header() const = 0;
};

class MemDFJK : public JK {
   protected:
    Options& options_;
    std::string name() override { return "MemDFJK"; }
    size_t memory_estimate() override;
    std::shared_ptr<DFHelper> dfh_;
    std::shared_ptr<BasisSet> auxiliary_;
    int df_ints_num_threads_;
    double condition_ = 1.0E-12;
    int max_nocc() const;
    bool C1() const override { return true; }
    void preiterations() override;
    void compute_JK() override;
    void postiterations() override;
    void common_init();

   public:
    MemDFJK(std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> auxiliary, Options& options);
    ~MemDFJK() override;
    void set_condition(double condition) { condition_ = condition; }
    void set_df_ints_num_threads(int val) { df_ints_num_threads_ = val; }
    void set_do_wK(bool do_wK) override;
    void print_header() const override;
    void set_omega_alpha(double alpha) override;
    void set_omega_beta(double beta) override;
    void set_wcombine(bool wcombine) override;
    void set_cutoff(double cutoff) override;
    std::shared_ptr<DFHelper> dfh() { return dfh_; }
};
}  // namespace psi
