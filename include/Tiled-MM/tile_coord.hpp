// Drop-in counterpart of reference src/Tiled-MM/tile_coord.hpp: (row, column) index of a tile in a tile grid.  Header-only here.
#pragma once

namespace gpu {

struct tile_coord {
    tile_coord() = default;
    tile_coord(int tile_idx_row, int tile_idx_col) : row_(tile_idx_row), col_(tile_idx_col) {}
    int row_index() const { return row_; }
    int col_index() const { return col_; }

private:
    int row_ = 0;
    int col_ = 0;
};

}  // namespace gpu
