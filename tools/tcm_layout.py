"""RBCM -> TCM re-layout of saved-activation / dZ tiles (host-side helper of the pipe-stage probe and test).

RBCM block (mlp_tc.cuh): [row half 2][chunk n][row 64][16 B]; TCM block: [chunk n][row 128][16 B]."""
import torch

PANEL = 16384
HIDDEN_SLOTS = 9
SAVED_PANELS = 1 + HIDDEN_SLOTS * 4 + 2
SAVED_TILE_BYTES = SAVED_PANELS * PANEL + (HIDDEN_SLOTS + 1) * 8 * 128 * 4
DZ_PANELS = HIDDEN_SLOTS * 4 + 3 + 1
DZ_TILE_BYTES = DZ_PANELS * PANEL
# (byte offset inside the tile, chunks) of every RBCM block
SAVED_BLOCKS = [((1 + 4 * i) * PANEL, 32) for i in range(HIDDEN_SLOTS)] + [((1 + 4 * HIDDEN_SLOTS) * PANEL, 16)]
DZ_BLOCKS = [(4 * i * PANEL, 32) for i in range(HIDDEN_SLOTS)] + [(4 * HIDDEN_SLOTS * PANEL, 24), ((4 * HIDDEN_SLOTS + 3) * PANEL, 8)]


def to_tcm(flat: torch.Tensor, n_tiles: int, tile_bytes: int, blocks, inverse: bool = False) -> torch.Tensor:
    """flat: uint8 tensor of n_tiles * tile_bytes bytes (1024-aligned start); returns a re-laid-out copy."""
    out = flat.clone()
    v_in, v_out = flat.view(n_tiles, tile_bytes), out.view(n_tiles, tile_bytes)
    for off, n in blocks:
        blk = v_in[:, off:off + n * 2048]
        if not inverse:
            v_out[:, off:off + n * 2048] = blk.reshape(n_tiles, 2, n, 64, 16).permute(0, 2, 1, 3, 4).reshape(n_tiles, -1)
        else:
            v_out[:, off:off + n * 2048] = blk.reshape(n_tiles, n, 2, 64, 16).permute(0, 2, 1, 3, 4).reshape(n_tiles, -1)
    return out
