/*
 * pybind11 export of B200MemDFJK, the analogue of the MemDFJK lines of psi4/src/export_fock.cc:193-194.
 * Built as its own extension module next to psi4.core (it only needs psi4's headers and core.so's symbols):
 *     import psi4, b200jk_psi4
 *     jk = b200jk_psi4.B200MemDFJK(wfn.basisset(), wfn.get_basisset("DF_BASIS_SCF"), ngpu=8)
 *     psi4.energy('scf', jk=jk)            # proc.py:2043-2045
 */
#include "psi4/pybind11.h"

#include "psi4/libmints/basisset.h"
#include "psi4/liboptions/liboptions.h"
#include "psi4/libpsi4util/process.h"

#include "B200MemDFJK.h"

namespace py = pybind11;
using namespace psi;

PYBIND11_MODULE(b200jk_psi4, m) {
    py::module_::import("psi4.core");  // registers JK / MemDFJK / BasisSet first
    py::class_<B200MemDFJK, std::shared_ptr<B200MemDFJK>, MemDFJK>(m, "B200MemDFJK", "MEM_DF J/K on B200 GPUs")
        .def(py::init([](std::shared_ptr<BasisSet> primary, std::shared_ptr<BasisSet> aux, int ngpu, bool release_host) {
                 // configured like JK::build_JK configures a MEM_DF object (jk.cc:143-148, _set_dfjk_options :58-68)
                 return B200MemDFJK::build(primary, aux, Process::environment.options, ngpu, release_host);
             }),
             py::arg("primary"), py::arg("auxiliary"), py::arg("ngpu") = 1, py::arg("release_host") = true)
        .def("last_stats", [](const B200MemDFJK& jk) {
            b200jk_stats s = jk.last_stats();
            py::dict d;
            d["ms_total"] = s.ms_total;
            d["ms_j"] = s.ms_j;
            d["ms_half"] = s.ms_half;
            d["ms_kgemm"] = s.ms_kgemm;
            d["ms_allreduce"] = s.ms_allreduce;
            d["ms_h2d"] = s.ms_h2d;
            d["ms_d2h"] = s.ms_d2h;
            d["launches"] = s.launches;
            return d;
        });
}
