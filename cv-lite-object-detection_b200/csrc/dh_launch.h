// Host helpers shared by the launchers in encode.cu / loss.cu.
#pragma once
#include "dh_tile.cuh"

namespace dh {
// Fill tile_begin / n_tiles / fast divisors for tt.maps[0..n_maps) and pick rows_per_tile so that one
// tile is about `tile_bytes`.  Returns the shared-memory bytes one stage buffer needs.
int finish_table(TileTable& tt, int ch, int batch, int tile_bytes);
}  // namespace dh
