// swd_window.cuh — sliding-window bookkeeping kernels (guessing.py:204-227, osd.py:170-179):
// window syndrome extraction, commit of the first F rounds with the sparse syndrome update
// new_det = det + e_hat @ chk.T (the reference does a dense float GEMM), failure counters.
#pragma once
#include "swd_device.cuh"

__global__ void accumulate_count_kernel(const int *counters, u64 *stats) { stats[5] += (u64)counters[0]; }

__global__ void window_extract_kernel(const u8 *__restrict__ det, long long B, int num_det, int row0, int m, u8 *__restrict__ synd) {
    const long long total = B * m;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / m; const int r = (int)(i - b * m);
        synd[i] = det[b * num_det + row0 + r];
    }
}

// one warp per shot: lanes scan 32 committed columns at a time (coalesced), every set bit's
// detector / observable rows are XORed by the first lanes (rows of one column are distinct).
__global__ void window_commit_kernel(const u8 *__restrict__ corr, long long B, int n_win, int col0, int ncommit,
                                     const int *__restrict__ chk_cp, const int *__restrict__ chk_ri, int num_det, u8 *det,
                                     const int *__restrict__ obs_cp, const int *__restrict__ obs_ri, int num_obs, u8 *obs) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long b = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += (long long)gridDim.x * wpb) {
        const u8 *row = corr + b * n_win;
        for (int base = 0; base < ncommit; base += 32) {
            const int j = base + lane;
            u32 bits = __ballot_sync(FULLMASK, j < ncommit && row[j] != 0);
            while (bits) {
                const int k = __ffs(bits) - 1; bits &= bits - 1;
                const int c = col0 + base + k;
                for (int e = chk_cp[c] + lane; e < chk_cp[c + 1]; e += 32) det[b * num_det + chk_ri[e]] ^= 1;
                if (obs && num_obs > 0) for (int e = obs_cp[c] + lane; e < obs_cp[c + 1]; e += 32) obs[b * num_obs + obs_ri[e]] ^= 1;
                __syncwarp();
            }
        }
    }
}

__global__ void window_count_kernel(const u8 *__restrict__ det, int num_det, const u8 *__restrict__ obs, int num_obs, long long B, u64 *out2) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    u64 nflag = 0, nfail = 0;
    for (long long b = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += (long long)gridDim.x * wpb) {
        int f = 0, l = 0;
        for (int r = lane; r < num_det; r += 32) f |= det[b * num_det + r];
        if (obs) for (int r = lane; r < num_obs; r += 32) l |= obs[b * num_obs + r];
        f = __any_sync(FULLMASK, f); l = __any_sync(FULLMASK, l);
        nflag += f ? 1 : 0; nfail += (f || l) ? 1 : 0;
    }
    if (lane == 0) { if (nflag) atomicAdd(&out2[0], nflag); if (nfail) atomicAdd(&out2[1], nfail); }
}

// ----------------------------------------------------------------------------------------------
// DEM sampling on the device: what stim's CompiledDemSampler.sample draws (guessing.py:129-130,
// build_circuit.py:271-288) - an independent Bernoulli(prior[c]) per DEM column, det = chk . e,
// obs = obs_mat . e (mod 2) - so that large runs never touch the host.
// Counter-based generator: Philox4x32-10 keyed by the seed, counter = (shot, column block), four
// columns per call; a column fires when its 32-bit draw is below floor(p * 2^32).  One warp per
// shot; detector / observable parities are accumulated in shared-memory bit words.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void window_sample_kernel(const u32 *__restrict__ thr, int num_col, const int *__restrict__ chk_cp, const int *__restrict__ chk_ri,
                                     int num_det, const int *__restrict__ obs_cp, const int *__restrict__ obs_ri, int num_obs,
                                     unsigned long long seed, long long shot0, long long B, u8 *__restrict__ det, u8 *__restrict__ obs,
                                     u8 *__restrict__ err) {
    extern __shared__ u32 s_bits[];                           // per warp: ceil(num_det/32) + ceil(num_obs/32) words
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int wd = (num_det + 31) >> 5, wo = (num_obs + 31) >> 5;
    u32 *dbits = s_bits + (size_t)wid * (wd + wo), *obits = dbits + wd;
    const int nblk = (num_col + 3) >> 2;
    for (long long b = (long long)blockIdx.x * wpb + wid; b < B; b += (long long)gridDim.x * wpb) {
        for (int i = lane; i < wd + wo; i += 32) dbits[i] = 0;
        __syncwarp();
        const unsigned long long shot = (unsigned long long)(shot0 + b);
        for (int blk = lane; blk < nblk; blk += 32) {
            u32 r[4];
            philox4x32_10((u32)blk, 0u, (u32)shot, (u32)(shot >> 32), (u32)seed, (u32)(seed >> 32), r);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = blk * 4 + q;
                if (c >= num_col) break;
                const bool fire = r[q] < thr[c];
                if (err) err[b * num_col + c] = (u8)fire;
                if (fire) {
                    for (int e = chk_cp[c]; e < chk_cp[c + 1]; e++) { const int row = chk_ri[e]; atomicXor(&dbits[row >> 5], 1u << (row & 31)); }
                    if (num_obs > 0) for (int e = obs_cp[c]; e < obs_cp[c + 1]; e++) { const int row = obs_ri[e]; atomicXor(&obits[row >> 5], 1u << (row & 31)); }
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < num_det; i += 32) det[b * num_det + i] = (u8)((dbits[i >> 5] >> (i & 31)) & 1u);
        if (obs) for (int i = lane; i < num_obs; i += 32) obs[b * num_obs + i] = (u8)((obits[i >> 5] >> (i & 31)) & 1u);
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------
// bit-packed shot I/O (SURVEY 7 step 1 / 8(b) "or bit-packed"): row b of `packed` holds ceil(nbits / 64) 64-bit words,
// bit j of the row = (word[j >> 6] >> (j & 63)) & 1 (numpy: packbits(bitorder="little") viewed as uint64).
// One warp per 1024 bits of a row: 32 coalesced 32-byte reads + ballots, one coalesced 128-byte store (and the reverse).
// ----------------------------------------------------------------------------------------------
__global__ void pack_bits_kernel(const u8 *__restrict__ bytes, long long B, int nbits, int words64, u32 *__restrict__ packed) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int segs = (words64 * 2 + 31) / 32;                   // segments of 32 32-bit words per row
    if (warp >= B * segs) return;
    const long long b = warp / segs; const int seg = (int)(warp % segs);
    const u8 *row = bytes + b * nbits;
    u32 mine = 0;
#pragma unroll 4
    for (int k = 0; k < 32; k++) {
        const int j = (seg * 32 + k) * 32 + lane;
        const u32 w = __ballot_sync(FULLMASK, j < nbits && row[j] != 0);
        if (lane == k) mine = w;
    }
    const int wi = seg * 32 + lane;
    if (wi < words64 * 2) packed[b * words64 * 2 + wi] = mine;
}

__global__ void unpack_bits_kernel(const u32 *__restrict__ packed, long long B, int nbits, int words64, u8 *__restrict__ bytes) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int segs = (words64 * 2 + 31) / 32;
    if (warp >= B * segs) return;
    const long long b = warp / segs; const int seg = (int)(warp % segs);
    const int wi = seg * 32 + lane;
    const u32 mine = (wi < words64 * 2) ? packed[b * words64 * 2 + wi] : 0u;
    u8 *row = bytes + b * nbits;
#pragma unroll 4
    for (int k = 0; k < 32; k++) {
        const u32 w = __shfl_sync(FULLMASK, mine, k);
        const int j = (seg * 32 + k) * 32 + lane;
        if (j < nbits) row[j] = (u8)((w >> lane) & 1u);
    }
}
