// HARNESS STAND-IN for psi4/libmints/matrix.h:69- (C1 matrices only: one irrep, contiguous row-major block, which is
// what Matrix::get_pointer() / pointer()[0] hand out in the reference, matrix.cc:3586-3593, matrix.h:553).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "psi4/libmints/dimension.h"
namespace psi {
class Matrix {
    std::string name_;
    Dimension rows_, cols_;
    std::vector<double> data_;
    std::vector<double*> rowptr_;

   public:
    Matrix(const std::string& name, int rows, int cols) : name_(name), rows_(1, rows), cols_(1, cols) { alloc(); }
    Matrix(const std::string& name, const Dimension& rows, const Dimension& cols, int /*symmetry*/ = 0)
        : name_(name), rows_(rows), cols_(cols) {
        alloc();
    }
    Matrix(int rows, int cols) : Matrix("", rows, cols) {}
    void alloc() {
        data_.assign(static_cast<size_t>(rows_[0]) * cols_[0] + 1, 0.0);
        rowptr_.resize(rows_[0] ? rows_[0] : 1);
        for (int r = 0; r < rows_[0]; r++) rowptr_[r] = data_.data() + static_cast<size_t>(r) * cols_[0];
        if (!rows_[0]) rowptr_[0] = data_.data();
    }
    double* get_pointer(const int& = 0) const { return const_cast<double*>(data_.data()); }
    double** pointer(const int& = 0) { return rowptr_.data(); }
    const Dimension& rowspi() const { return rows_; }
    const Dimension& colspi() const { return cols_; }
    int nirrep() const { return 1; }
    int symmetry() const { return 0; }
    void zero() { std::fill(data_.begin(), data_.end(), 0.0); }
    void hermitivitize() {  // matrix.cc: (A + A^T)/2
        const int n = rows_[0];
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++) {
                double v = 0.5 * (data_[static_cast<size_t>(i) * n + j] + data_[static_cast<size_t>(j) * n + i]);
                data_[static_cast<size_t>(i) * n + j] = data_[static_cast<size_t>(j) * n + i] = v;
            }
    }
};
using SharedMatrix = std::shared_ptr<Matrix>;
}  // namespace psi
