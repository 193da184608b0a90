// k_em_psum: the class-owner-centric EM kernel (variant 5). Data layout shared by the packer (prep_psum.cu) and the kernel (em_psum.cu).
//
// Ownership: the participating rows, in the index's locality order, are cut into one contiguous range per CTA (= per SM); a class belongs
// to the CTA that owns its MEDIAN member, so the classes of a wide module spread over the CTAs that hold its rows. A CTA keeps in SHARED
// MEMORY, for the whole kernel: theta of its rows and of every other row its classes touch ("halo rows"), q of its classes, and the partial
// row sums it is about to hand over. Both phases then run over the CTA's own classes only:
//   E      q_c = R_c / sum_{t in c} theta_t                                  class-major tiles, gathers from shared memory
//   M      S_t^(b) = sum_{c owned by b, c contains t} q_c                      row-major over the same members, gathers from shared memory
//   U      theta_t' = (Rs_t + theta_t * sum_b S_t^(b)) / A_t                   by the row's owner, contributions added in CTA order
// Across CTAs only per-row values travel, through tagged 16-byte slots (ll_store / ll_load): theta_t to the CTAs whose classes contain t,
// and S_t^(b) back to the owner of t. q never leaves its CTA, a hub row costs its owner one slot per contributing CTA, and a class of 999
// members costs its owner 999 shared-memory gathers. Every index is a 16-bit shared-memory slot.
#pragma once
#include "common.cuh"

constexpr int PS_MAX_SLOT = 65534;         // 0xFFFF marks an unused lane of an M slice

// ---- E side -------------------------------------------------------------------------------------------------------------------------
// cardinality 2: a tile holds 128 classes, lane l carries classes l, 32 + l, 64 + l, 96 + l as 4 x (2 u16) = 16 bytes;
// cardinality 3, 4: 64 classes, lane l carries classes l, 32 + l as 2 x (4 u16) = 16 bytes (the 4th entry of a 3-member class is the zero slot);
// otherwise G = 1 << e_lgG(k) lanes share a class (32 / G classes per tile) and a lane's members come in chunks of 4 u16 (8 bytes):
// chunk c of all 32 lanes is one 256-byte block. Entries beyond the cardinality point at the zero-theta slot.
__host__ __device__ __forceinline__ int ps_steps4(int k) { return (e_steps(k) + 3) >> 2; }
__host__ __device__ __forceinline__ int ps_tile_u16(int k) { return k <= 4 ? 256 : 128 * ps_steps4(k); }      // 16-bit words of index data per tile
// Tiles of classes that share lanes (cardinality > 16) cover up to PS_NB consecutive row blocks of their cell, which the warp streams as one
// run of chunks: the latency of the first load is paid once per tile, not once per 1 - 8 classes.
constexpr int PS_NB = 8;
__host__ __device__ __forceinline__ int ps_blocks_per_tile(int k) { return e_lgG(k) > 0 ? PS_NB : 1; }
// a resident copy of a tile (shared memory) carries its read counts right behind the index data

// ---- M side -------------------------------------------------------------------------------------------------------------------------
// The rows a CTA's classes touch (own rows and halo rows), sorted by their number of local entries, longest first.
// A row's destination is one 32-bit word: its slot in the CTA's partial-sum array when the CTA owns the row, or PS_REMOTE | the
// partial-sum slot of the row's owner | the owner's rank << 28 for a halo row; PS_NONE = unused lane.
// rows with more than M_LONG entries: groups of <= M_GROUP_ROWS rows, one warp per group: header = two 32-bit words {length, destination}
// per row (padded to 16 bytes), then the rows' entries back to back, each row padded to an even number of entries;
// the rest: slices of 32 rows: header = 32 destinations (128 bytes), then chunks of 4 entries per lane (chunk c of all lanes = 256
// bytes), padded with the zero-q slot up to the longest row of the slice.
constexpr uint32_t PS_REMOTE = 0x80000000u, PS_NONE = 0xFFFFFFFFu;
__host__ __device__ __forceinline__ int ps_slice_u16(int len) { return 64 + 128 * ((len + 3) >> 2); }
__host__ __device__ __forceinline__ int ps_group_hdr_u16(int rows) { return ((2 * rows + 3) & ~3) * 2; }

constexpr int PS_STG_BYTES = 1152 * EM_WARPS;     // one staging buffer per warp (em_psum.cu)
struct PsPlan { int off_et, off_mi, off_theta, off_q, off_Q, off_in, off_stg, off_cache, total; };
// nin: partial sums the CTA receives per iteration (staged in shared memory before its rows are updated)
__host__ __device__ __forceinline__ PsPlan ps_smem_plan(int desc_smem, int n_et, int n_mi, int nrows, int nhr, int ncls, int nin, int stage)
{
    PsPlan p;
    p.off_et = 0;
    p.off_mi = p.off_et + (desc_smem ? n_et * 16 : 0);
    p.off_theta = p.off_mi + (desc_smem ? n_mi * 16 : 0);
    p.off_q = p.off_theta + (((nrows + nhr + 1) * 8 + 15) & ~15);
    p.off_Q = p.off_q + (((ncls + 1) * 8 + 15) & ~15);
    p.off_in = p.off_Q + ((nrows * 8 + 15) & ~15);
    p.off_stg = p.off_in + ((nin * 8 + 15) & ~15);
    p.off_cache = p.off_stg + (stage ? PS_STG_BYTES : 0);
    p.total = p.off_cache;
    return p;
}
