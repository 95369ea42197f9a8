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
