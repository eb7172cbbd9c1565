/*
 * probe_types.cc -- TEST INFRASTRUCTURE.  Compiled (-fsyntax-only) twice by tests/test_glue_harness.py: against the
 * reference's REAL headers (with the declaration-only stand-ins of glue/compile_check for Eigen / Libint2) and against the
 * harness stand-ins of glue/harness/include.  Every member, type and virtual signature the glue relies on is asserted
 * here, so the harness the glue is EXECUTED against cannot drift from the classes it will meet inside psi4.
 */
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "psi4/lib3index/dfhelper.h"
#include "psi4/libfock/jk.h"
#include "psi4/libmints/matrix.h"

namespace psi {

struct ProbeDFHelper : public DFHelper {
    ProbeDFHelper(std::shared_ptr<BasisSet> p, std::shared_ptr<BasisSet> a) : DFHelper(p, a) {}
    void touch() {
        static_assert(std::is_same<decltype(nbf_), size_t>::value, "nbf_");
        static_assert(std::is_same<decltype(naux_), size_t>::value, "naux_");
        static_assert(std::is_same<decltype(AO_core_), bool>::value, "AO_core_");
        static_assert(std::is_same<decltype(do_wK_), bool>::value, "do_wK_");
        static_assert(std::is_same<decltype(sparsity_prepared_), bool>::value, "sparsity_prepared_");
        static_assert(std::is_same<decltype(Ppq_), std::unique_ptr<double[]>>::value, "Ppq_");
        static_assert(std::is_same<decltype(m1Ppq_), std::unique_ptr<double[]>>::value, "m1Ppq_");
        static_assert(std::is_same<decltype(wPpq_), std::unique_ptr<double[]>>::value, "wPpq_");
        static_assert(std::is_same<decltype(small_skips_), std::vector<size_t>>::value, "small_skips_");
        static_assert(std::is_same<decltype(big_skips_), std::vector<size_t>>::value, "big_skips_");
        static_assert(std::is_same<decltype(schwarz_fun_index_), std::vector<size_t>>::value, "schwarz_fun_index_");
        static_assert(std::is_same<decltype(symm_big_skips_), std::vector<size_t>>::value, "symm_big_skips_");
        prepare_sparsity();
        initialize();
        set_schwarz_cutoff(1e-12);
        set_fitting_condition(1e-10);
        set_do_wK(false);
        set_omega(0.0);
        set_wcombine(false);
        (void)ao_sparsity();
        (void)get_AO_core();
    }
};

struct ProbeMemDFJK : public MemDFJK {
    ProbeMemDFJK(std::shared_ptr<BasisSet> p, std::shared_ptr<BasisSet> a, Options& o) : MemDFJK(p, a, o) {}
    // the virtuals the glue overrides, with the reference's exact signatures and access
    std::string name() override { return "probe"; }
    void preiterations() override { MemDFJK::preiterations(); }
    void compute_JK() override {}
    void postiterations() override {}
    void print_header() const override { MemDFJK::print_header(); }
    void touch(Options& options) {
        static_assert(std::is_same<decltype(dfh_), std::shared_ptr<DFHelper>>::value, "dfh_");
        static_assert(std::is_same<decltype(condition_), double>::value, "condition_");
        static_assert(std::is_same<decltype(C_left_ao_), std::vector<SharedMatrix>>::value, "C_left_ao_");
        static_assert(std::is_same<decltype(C_right_ao_), std::vector<SharedMatrix>>::value, "C_right_ao_");
        static_assert(std::is_same<decltype(D_ao_), std::vector<SharedMatrix>>::value, "D_ao_");
        static_assert(std::is_same<decltype(J_ao_), std::vector<SharedMatrix>>::value, "J_ao_");
        static_assert(std::is_same<decltype(K_ao_), std::vector<SharedMatrix>>::value, "K_ao_");
        static_assert(std::is_same<decltype(wK_ao_), std::vector<SharedMatrix>>::value, "wK_ao_");
        static_assert(std::is_same<decltype(do_J_), bool>::value, "do_J_");
        static_assert(std::is_same<decltype(do_K_), bool>::value, "do_K_");
        static_assert(std::is_same<decltype(do_wK_), bool>::value, "do_wK_");
        static_assert(std::is_same<decltype(lr_symmetric_), bool>::value, "lr_symmetric_");
        static_assert(std::is_same<decltype(print_), int>::value, "print_");
        static_assert(std::is_same<decltype(cutoff_), double>::value, "cutoff_");
        static_assert(std::is_same<decltype(omega_), double>::value, "omega_");
        set_wcombine(false);
        set_cutoff(1e-12);
        set_print(1);
        set_debug(0);
        set_bench(0);
        set_condition(1e-10);
        set_df_ints_num_threads(1);
        set_do_J(true);
        set_do_K(true);
        set_do_wK(false);
        set_omega(0.0);
        (void)options;
        SharedMatrix m = C_left_ao_.empty() ? SharedMatrix() : C_left_ao_[0];
        if (m) {
            double* p = m->get_pointer();
            int r = m->rowspi()[0], c = m->colspi()[0];
            (void)p;
            (void)r;
            (void)c;
        }
    }
};

}  // namespace psi
