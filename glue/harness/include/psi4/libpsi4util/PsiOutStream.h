// HARNESS STAND-IN for psi4/libpsi4util/PsiOutStream.h: Printf into a string the test can read back.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <string>
namespace psi {
class PsiOutStream {
    std::string text_;

   public:
    void Printf(const char* fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        text_ += buf;
    }
    const std::string& text() const { return text_; }
    void clear() { text_.clear(); }
};
extern std::shared_ptr<PsiOutStream> outfile;
}  // namespace psi
