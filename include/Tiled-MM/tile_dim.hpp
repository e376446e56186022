// Drop-in counterpart of reference src/Tiled-MM/tile_dim.hpp: rows x cols of one tile.  Same accessors; the element
// count is kept in 64 bits as well (the reference's int product overflows for tiles >= 2^31 elements, tile_dim.cpp:5-8),
// and a default-constructed tile is 0 x 0 instead of indeterminate.
#pragma once
#include <cstddef>

namespace gpu {

class tile_dim {
public:
    tile_dim() = default;
    tile_dim(int n_rows, int n_cols) : n_rows_(n_rows), n_cols_(n_cols) {}

    int rows() const { return n_rows_; }
    int cols() const { return n_cols_; }
    void set_rows(int n_rows) { n_rows_ = n_rows; }
    void set_cols(int n_cols) { n_cols_ = n_cols; }
    int size() const { return n_rows_ * n_cols_; }
    std::size_t size64() const { return (std::size_t)n_rows_ * (std::size_t)n_cols_; }

private:
    int n_rows_ = 0;
    int n_cols_ = 0;
};

}  // namespace gpu
