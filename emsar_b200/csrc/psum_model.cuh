// The packed model k_em_psum runs on (see psum.cuh for the layout). Plain pointers into device memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct PsModel {
    int32_t P, B, block0;          // rows; CTAs of this launch; first virtual CTA of this device (0 on a single GPU)
    int64_t C_a, nnz_a;
    int32_t *blk_row0, *blk_cls0, *blk_etile0, *blk_mitem0, *blk_hr0;     // [Bt + 1] per virtual CTA
    int32_t *blk_desc_smem;        // [Bt] 1: the CTA keeps its tile / item descriptors in shared memory
    int4 *e_tiles;                 // {first local class, classes, data offset (16-byte units; bit 30 of w set: offset inside the CTA's resident cache), steps | lgG << 12 | resident << 30}
    int4 *m_items;                 // {-, rows, data offset (as above), length | resident << 29 | group << 30}
    int32_t *e_src, *m_src;        // global data offset of every tile / item (the resident cache is filled from it)
    int32_t *blk_res16;            // [Bt] 16-byte units of the CTA's resident cache
    const unsigned char *e_data, *m_data;
    const uint32_t *e_R;           // [C_a] read counts in compact class order
    int32_t *halo_rows;            // [n_inc] global row of every halo slot, grouped by CTA (blk_hr0)
    int32_t *halo_tgt;             // [n_inc] partial-sum slot the CTA's contribution to that row goes to | rank of the row's owner << 28
    int32_t *inc_off;              // [P + 1] partial-sum slots of row p: inc_off[p] .. inc_off[p + 1], ordered by contributing CTA
    double2 *row_RsA;              // [P] {Rs, A}
    double *theta;                 // [P]
    int32_t *row_mask;             // [P] ranks that hold a CTA contributing to (= reading theta of) row p, one bit per rank
    // exchange slots (16 bytes, tagged): every rank holds the three arrays at the same offsets of its window; win[r] is rank r's window as
    // mapped into this device (NVLink peer memory), win[rank] the local one. On a single GPU: one rank, the window is the sample's slot buffer.
    unsigned char *win[8];
    long long th_off;              // [P + 1]      theta of every row that some other CTA reads (written by the row's owner into every reader rank)
    long long part_off;            // [n_inc + 1]  partial row sums (written by the contributor into the owner's rank)
    long long dm_off;              // [2 * Bt]     convergence measure per CTA, alternating by iteration parity (written into every rank)
    int32_t rank, nranks;
    int32_t stage;                 // bit 0: E tiles, bit 1: M items that are not resident travel through the per-warp staging buffers (cp.async look-ahead)
    int32_t Bt;                    // virtual CTAs over all devices (= B on a single GPU)
    int32_t n_inc;
    int32_t smem_bytes;
};

