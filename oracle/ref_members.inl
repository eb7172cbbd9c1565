    // ---- state of DFHelper that the sliced functions read (names and types of lib3index/dfhelper.h:333-452) ----
    size_t nbf_ = 0, naux_ = 0, nthreads_ = 1, memory_ = 0, max_nocc_ = 0;
    int debug_ = 0;
    bool AO_core_ = true, direct_ = false, wcombine_ = false, do_wK_ = false, hold_met_ = true;
    std::string method_ = "STORE";
    std::unique_ptr<double[]> Ppq_, m1Ppq_, wPpq_;
    std::vector<size_t> small_skips_, big_skips_, symm_small_skips_, symm_ignored_columns_, symm_big_skips_;
    std::vector<size_t> schwarz_fun_index_, schwarz_shell_mask_, Qshell_aggs_;
    size_t Qshells_ = 0, pshells_ = 0;
    double cutoff_ = 1e-12;
    bool sparsity_prepared_ = false;
    std::vector<std::string> AO_names_ = {"", ""};
    std::map<std::string, std::tuple<std::string, std::string>> files_;
    // ---- the disk / metric-file side is never reached by the in-core MEM_DF path: abort loudly if it were ----
    [[noreturn]] static void unreachable(const char* what) { throw std::logic_error(std::string("ref shim: ") + what); }
    void stream_check(std::string, std::string) { unreachable("stream_check (disk path)"); }
    void grab_AO(size_t, size_t, double*) { unreachable("grab_AO (disk path)"); }
    std::string return_metfile(double) { unreachable("return_metfile"); }
    void get_tensor_(std::string, double*, size_t, size_t, size_t, size_t) { unreachable("get_tensor_"); }
    double* metric_prep_core(double) { unreachable("metric_prep_core"); }
