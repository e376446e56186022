// Drop-in counterpart of reference src/Tiled-MM/tiled_matrix.{hpp,cpp}: a non-owning tiling of a column-major host matrix - tile
// size clamped to the matrix, tile count by ceiling, the last tile of a row / column of tiles is the remainder (tiled_matrix.cpp:8-23,
// 31-39, 62-80).  The scheduler of this library does not use it (it plans panels, k-chunks and column blocks, csrc/tmm_plan.cpp); it
// is kept, header-only, for callers that include it.  Difference: the element offset of a tile is also available in 64 bits
// (tile_offset64) - the reference's int product overflows from 2^31 elements on - and zero-sized matrices are legal.
#pragma once
#include "tile_coord.hpp"
#include "tile_dim.hpp"

#include <algorithm>
#include <cstddef>

namespace gpu {

template <typename Scalar>
class tiled_matrix {
public:
    tiled_matrix(Scalar* host_ptr, int rows, int cols, int ld, tile_dim d)
        : ptr_(host_ptr), rows_(rows), cols_(cols), ld_(ld), tile_(std::min(d.rows(), rows), std::min(d.cols(), cols)) {
        tiles_row_ = tile_.rows() > 0 ? (rows_ + tile_.rows() - 1) / tile_.rows() : 0;
        tiles_col_ = tile_.cols() > 0 ? (cols_ + tile_.cols() - 1) / tile_.cols() : 0;
    }

    // nominal tile, and the actual extent of one tile (the last one in each direction may be shorter)
    tile_dim tile_dimensions() { return tile_; }
    tile_dim tile_dimensions(tile_coord t) { return tile_dim(extent(rows_, tile_.rows(), tiles_row_, t.row_index()), extent(cols_, tile_.cols(), tiles_col_, t.col_index())); }

    int rows() { return rows_; }
    int cols() { return cols_; }
    int leading_dim() { return ld_; }
    Scalar* data() { return ptr_; }
    int num_tiles_row() { return tiles_row_; }
    int num_tiles_col() { return tiles_col_; }

    // element offset of a tile's first entry: column-major, col * tile_cols * ld + row * tile_rows
    std::size_t tile_offset64(tile_coord t) {
        return (std::size_t)t.col_index() * (std::size_t)tile_.cols() * (std::size_t)ld_ + (std::size_t)t.row_index() * (std::size_t)tile_.rows();
    }
    int tile_offset(tile_coord t) { return (int)tile_offset64(t); }
    Scalar* tile_data(tile_coord t) { return ptr_ + tile_offset64(t); }

private:
    static int extent(int dim, int tile, int n_tiles, int id) {
        if (tile <= 0 || id < 0 || id >= n_tiles) return 0;
        return id + 1 < n_tiles ? tile : dim - tile * (n_tiles - 1);
    }
    Scalar* ptr_;
    int rows_, cols_, ld_;
    tile_dim tile_;
    int tiles_row_ = 0, tiles_col_ = 0;
};

}  // namespace gpu
