/*
 * i3d_recolor.cuh — voxel recolouring on the device.
 *
 * Replaces Intrinsic3D::recomputeColors (libintrinsic3d/src/refinement/intrinsic3d.cpp:381-409), i.e.
 * SDFColorization::add for every frame + SDFColorization::compute (src/sdf/colorization.cpp:113-189), with
 * computeObservation / computeWeight / filter / computeColor (:215-370) and interpolateRGB (src/rgbd/processing.cpp:236-302).
 *
 * One thread per voxel that has a forward-difference normal.  The frame scan is the one of k_select_obs (same float pipeline,
 * same conservative per-warp frame culling); the best K (weight, frame) keys stay in registers.  The reference keeps a
 * std::vector of observations per voxel (N x F VertexObservation objects across the F add() calls); here nothing is stored:
 * the <= K winning frames are re-projected at the end and their colours fetched bilinearly.
 *
 * Float summation order is the reference's: when the top-K filter runs (more than K observations) the colours are summed in
 * ascending (weight, frame) order (the order std::sort leaves them in; ties broken by frame id, the canonical choice of
 * oracle.cpp), otherwise (K == 0 or at most K observations) in frame order.
 */
#pragma once
#include "i3d_kernels.cuh"

namespace i3d
{

// interpolate<unsigned char> (src/rgbd/processing.cpp:236-291): bilinear on an interleaved B,G,R image, out-of-image taps dropped
__device__ __forceinline__ unsigned char interp_u8(const uint8_t* __restrict__ img, int w, int h, float x, float y, int channel)
{
    const float fx0 = floorf(x), fy0 = floorf(y);
    const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
    const int x1 = x0 + 1, y1 = y0 + 1;
    float x1w = FS(x, fx0), y1w = FS(y, fy0);
    float x0w = FS(1.0f, x1w), y0w = FS(1.0f, y1w);
    if (x0 < 0 || x0 >= w) x0w = 0.0f;
    if (x1 < 0 || x1 >= w) x1w = 0.0f;
    if (y0 < 0 || y0 >= h) y0w = 0.0f;
    if (y1 < 0 || y1 >= h) y1w = 0.0f;
    const float w00 = FM(x0w, y0w), w10 = FM(x1w, y0w), w01 = FM(x0w, y1w), w11 = FM(x1w, y1w);
    const float sum_w = FA(FA(FA(w00, w10), w01), w11);
    float sum = 0.0f;
    if (w00 > 0.0f) sum = FA(sum, FM(static_cast<float>(img[(static_cast<size_t>(y0) * w + x0) * 3 + channel]), w00));
    if (w01 > 0.0f) sum = FA(sum, FM(static_cast<float>(img[(static_cast<size_t>(y1) * w + x0) * 3 + channel]), w01));
    if (w10 > 0.0f) sum = FA(sum, FM(static_cast<float>(img[(static_cast<size_t>(y0) * w + x1) * 3 + channel]), w10));
    if (w11 > 0.0f) sum = FA(sum, FM(static_cast<float>(img[(static_cast<size_t>(y1) * w + x1) * 3 + channel]), w11));
    if (!(sum_w > 0.0f)) return 0;
    return static_cast<unsigned char>(__float2int_rz(FD(sum, sum_w)));
}

// K == 0: no filter (all observations, frame order).  counts[0] += voxels recoloured, counts[1] += observations with weight > 0.
template <int KMAX>
__global__ void __launch_bounds__(kThreads)
k_recolor(GridView g, FrameView fr, const uint8_t* __restrict__ bgr /* [F][H][W][3] */, const float* __restrict__ Rt, SelectCam cam, CullView cull, int K,
          uchar4* __restrict__ rgb_out, unsigned long long* __restrict__ counts)
{
    extern __shared__ float s_rt[];     // [F][12]
    for (int i = threadIdx.x; i < 12 * fr.F; i += blockDim.x) s_rt[i] = Rt[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    float nrm[3] = {0.0f, 0.0f, 0.0f};
    float pt[3] = {0.0f, 0.0f, 0.0f};
    bool in_range = false;
    if (v < g.n && surface_normal_f(g, v, nrm))             // add(): voxels without a normal collect nothing
    {
        in_range = true;
        const float s = static_cast<float>(g.sdf[v]);
        pt[0] = FS(FM(static_cast<float>(g.x[v]), g.voxel_size), FM(nrm[0], s));
        pt[1] = FS(FM(static_cast<float>(g.y[v]), g.voxel_size), FM(nrm[1], s));
        pt[2] = FS(FM(static_cast<float>(g.z[v]), g.voxel_size), FM(nrm[2], s));
    }
    if (__ballot_sync(0xffffffffu, in_range) == 0u) return;
    // ---- candidate-frame mask of the warp's cluster (see k_select_obs / frame_may_see)
    const int nwords = (fr.F + 31) / 32;
    __shared__ unsigned s_mask[kThreads / 32][kCullMaxWords];
    unsigned* wmask = s_mask[threadIdx.x >> 5];
    const bool culling = cull.enabled && nwords <= kCullMaxWords;
    if (culling)
    {
        const float big = 3.0e38f;
        float lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = in_range ? pt[k] : big; hi[k] = in_range ? pt[k] : -big; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        const float c[3] = {0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2])};
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        const float rad = 0.5f * sqrtf(dx * dx + dy * dy + dz * dz) * 1.001f + 1e-4f;
#pragma unroll 1
        for (int j = 0; j < nwords; ++j)
        {
            const int f = 32 * j + lane;
            const bool may = (f < fr.F) && frame_may_see(c, rad, s_rt + 12 * f, cam, cull, f, fr.W, fr.H);
            const unsigned m = __ballot_sync(0xffffffffu, may);
            if (lane == 0) wmask[j] = m;
        }
        __syncwarp();
    }
    const size_t img = static_cast<size_t>(fr.W) * fr.H;
    const float scale_color = FD(1.0f, 255.0f);
    unsigned long long best[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) best[k] = 0ull;
    int n_obs = 0;
    float c3[3] = {0.0f, 0.0f, 0.0f}, wsum = 0.0f;
    auto add_color = [&](int f, float wf) {
        float pix[2];
        observation_weight(pt, nrm, s_rt + 12 * f, cam, fr.depth + img * f, fr.W, fr.H, pix);     // same arithmetic => same sub-pixel position
        const uint8_t* cimg = bgr + img * f * 3;
        const float ws = FM(wf, scale_color);
        c3[0] = FA(c3[0], FM(static_cast<float>(interp_u8(cimg, fr.W, fr.H, pix[0], pix[1], 2)), ws));
        c3[1] = FA(c3[1], FM(static_cast<float>(interp_u8(cimg, fr.W, fr.H, pix[0], pix[1], 1)), ws));
        c3[2] = FA(c3[2], FM(static_cast<float>(interp_u8(cimg, fr.W, fr.H, pix[0], pix[1], 0)), ws));
        wsum = FA(wsum, wf);
    };
#pragma unroll 1
    for (int j = 0; j < nwords; ++j)
    {
        unsigned m = culling ? wmask[j] : 0xffffffffu;
#pragma unroll 1
        while (m)
        {
            const int f = 32 * j + __ffs(m) - 1;
            m &= m - 1;
            if (f >= fr.F) continue;
            const float wf = observation_weight(pt, nrm, s_rt + 12 * f, cam, fr.depth + img * f, fr.W, fr.H);
            if (wf > 0.0f && in_range)
            {
                ++n_obs;
                if (K == 0) { add_color(f, wf); continue; }
                unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(wf)) << 32) | static_cast<unsigned>(f + 1);
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                {
                    const unsigned long long hi2 = key > best[k] ? key : best[k];
                    const unsigned long long lo2 = key > best[k] ? best[k] : key;
                    best[k] = hi2; key = lo2;
                }
            }
        }
    }
    // per-warp totals -> two atomics
    {
        const unsigned long long col = __popc(__ballot_sync(0xffffffffu, n_obs > 0));
        unsigned long long tot = static_cast<unsigned long long>(n_obs);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0 && counts) { atomicAdd(counts, col); atomicAdd(counts + 1, tot); }
    }
    if (n_obs == 0) return;                                  // compute(): no observation => colour unchanged
    if (K > 0)
    {
        // order the kept observations the way computeColor will meet them, then one (non-unrolled) colour loop
        if (n_obs > K)
        {
            // the filter ran: ascending (weight, frame) among the K kept ones == `best` read backwards; re-key as (frame, weight)
#pragma unroll
            for (int k = 0; k < KMAX / 2; ++k) { const unsigned long long t = best[k]; best[k] = best[KMAX - 1 - k]; best[KMAX - 1 - k] = t; }
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
            {
                // after the reversal, entries beyond the best K sit in front: drop them (slot index KMAX-1-k was the rank)
                const bool keep = (KMAX - 1 - k) < K && best[k] != 0ull;
                best[k] = keep ? (((best[k] & 0xffffffffull) << 32) | (best[k] >> 32)) : ~0ull;
            }
        }
        else
        {
            // the filter returned early: frame order
#pragma unroll
            for (int k = 0; k < KMAX; ++k) best[k] = (best[k] == 0ull) ? ~0ull : (((best[k] & 0xffffffffull) << 32) | (best[k] >> 32));
#pragma unroll
            for (int i = 0; i < KMAX; ++i)
#pragma unroll
                for (int j = 0; j + 1 < KMAX - i; ++j)
                {
                    const unsigned long long lo = best[j] < best[j + 1] ? best[j] : best[j + 1];
                    const unsigned long long hi = best[j] < best[j + 1] ? best[j + 1] : best[j];
                    best[j] = lo; best[j + 1] = hi;
                }
        }
        // best[] now holds (frame + 1) << 32 | weight bits in summation order, ~0 = empty
        int sel_f[KMAX]; float sel_w[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
        {
            sel_f[k] = best[k] == ~0ull ? -1 : static_cast<int>(best[k] >> 32) - 1;
            sel_w[k] = __uint_as_float(static_cast<unsigned>(best[k] & 0xffffffffull));
        }
#pragma unroll 1
        for (int k = 0; k < KMAX; ++k)
            if (sel_f[k] >= 0) add_color(sel_f[k], sel_w[k]);
    }
    // computeColor (colorization.cpp:318-354): mean colour, cast<unsigned char> truncates
    if (wsum > 0.0f) { const float s = FD(255.0f, wsum); c3[0] = FM(c3[0], s); c3[1] = FM(c3[1], s); c3[2] = FM(c3[2], s); }
    uchar4 o4 = rgb_out[v];
    o4.x = static_cast<unsigned char>(__float2int_rz(c3[0])); o4.y = static_cast<unsigned char>(__float2int_rz(c3[1])); o4.z = static_cast<unsigned char>(__float2int_rz(c3[2]));
    rgb_out[v] = o4;
}

__global__ void k_interleave_rgb(int64_t n, const uchar4* __restrict__ rgb, uint8_t* __restrict__ out3)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const uchar4 c = rgb[i];
    out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
}

} // namespace i3d
