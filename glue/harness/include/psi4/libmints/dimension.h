// HARNESS STAND-IN for psi4/libmints/dimension.h (test infrastructure: lets glue/B200MemDFJK.cc be compiled, linked and
// RUN without a psi4 build).  Only what the glue and the harness JK touch: operator[] and n().
#pragma once
#include <vector>
namespace psi {
class Dimension {
    std::vector<int> blocks_;

   public:
    Dimension() = default;
    explicit Dimension(int nirrep, int v = 0) : blocks_(nirrep, v) {}
    int n() const { return static_cast<int>(blocks_.size()); }
    int& operator[](int i) { return blocks_[i]; }
    const int& operator[](int i) const { return blocks_[i]; }
    bool operator==(const Dimension& o) const { return blocks_ == o.blocks_; }
    bool operator!=(const Dimension& o) const { return blocks_ != o.blocks_; }
};
}  // namespace psi
