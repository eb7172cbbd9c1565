// HARNESS STAND-IN for psi4/liboptions/liboptions.h: the five accessors JK::build_JK's option handling uses
// (libfock/jk.cc:58-68, :143-148).
#pragma once
#include <map>
#include <string>
namespace psi {
class Data {
    std::string s_;
    double d_ = 0.0;
    bool changed_ = false;

   public:
    Data() = default;
    Data(const std::string& s, double d, bool changed) : s_(s), d_(d), changed_(changed) {}
    bool has_changed() const { return changed_; }
    const std::string& str() const { return s_; }
    double num() const { return d_; }
};
class Options {
    std::map<std::string, Data> kv_;

   public:
    Options() {
        // defaults of read_options.cc for the keys the DF-JK factory reads
        set_default("SCREENING", "CSAM", 0.0);
        set_default("INTS_TOLERANCE", "", 1.0e-12);
        set_default("PRINT", "", 1);
        set_default("DEBUG", "", 0);
        set_default("BENCH", "", 0);
        set_default("DF_FITTING_CONDITION", "", 1.0e-10);  // read_options.cc:1734
        set_default("DF_INTS_NUM_THREADS", "", 0);
        set_default("WCOMBINE", "", 0);
    }
    void set_default(const std::string& k, const std::string& s, double d) { kv_[k] = Data(s, d, false); }
    void set_str(const std::string& k, const std::string& s) { kv_[k] = Data(s, 0.0, true); }
    void set_double(const std::string& k, double d) { kv_[k] = Data("", d, true); }
    void set_int(const std::string& k, int v) { kv_[k] = Data("", v, true); }
    std::string get_str(const std::string& k) { return kv_[k].str(); }
    double get_double(const std::string& k) { return kv_[k].num(); }
    int get_int(const std::string& k) { return static_cast<int>(kv_[k].num()); }
    bool get_bool(const std::string& k) { return kv_[k].num() != 0.0; }
    Data& operator[](const std::string& k) { return kv_[k]; }
};
}  // namespace psi
