// K5 / K6 -- what follows the Hamming search inside one LinkStage closure (reference src/pipeline/link_stage.cpp:83-88),
// on the device, so that a batched submission returns only what the host still needs:
//
//   K5  ratio test + order-preserving compaction. The reference emits a match iff `best < 0.8 * second` in IEEE double
//       on distance = popcount * (1.0 / 486) (src/match/match_features.cpp:79,94) and emits in query order (:71-97).
//       k5_count_kernel evaluates exactly that comparison (individually rounded __dmul_rn, the same constants) per
//       ocb_top2 record and counts the survivors of every pair; the last CTA to finish turns the counts into
//       exclusive offsets; k5_compact_kernel writes the survivors of pair p, in query order, densely at offsets[p].
//       The host then copies 12 bytes per SURVIVOR instead of 8 bytes per query and only runs the reference's
//       std::sort (:100-101) on them. HBM/L2 bound: 8 B read twice + 12 B written per survivor.
//
//   K6  distort_keypoints / image_to_3d (reference src/distort/distort_keypoints.cpp:48-103): pixel -> unit ray for
//       both sides of every match of a batch of pairs, written as the [n][7] correspondence rows
//       (include/opencalibration/types/correspondence.hpp:8-13) K2/K3 consume. Every operation is an individually
//       rounded IEEE double operation in the order of host/distort_keypoints.cpp (the C++ mirror, which restates the
//       Eigen expressions and the TinySolver Levenberg-Marquardt undistortion), so device rays == host rays bit for
//       bit. One thread per (match, side); FP64 pipe / latency bound, ~20 operations per ray without distortion.
#include "ocb_internal.cuh"
#include "std_sort_replay.cuh"

#include <algorithm>

namespace ocb
{

// ----------------------------------------------------------------------------------------------------------
// K5
// ----------------------------------------------------------------------------------------------------------
constexpr int K5_THREADS = 256;

__device__ __forceinline__ bool ratio_test_keeps(const ocb_top2 r)
{
    // match_features.cpp:79: distance = count * (1.0 / DESCRIPTOR_BITS); :74-75: best / second start at +infinity
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    const double unit = 1.0 / OCB_DESCRIPTOR_BITS;
    const double best = r.best_d == OCB_DIST_INF ? inf : __dmul_rn((double)r.best_d, unit);
    const double second = r.second_d == OCB_DIST_INF ? inf : __dmul_rn((double)r.second_d, unit);
    return best < __dmul_rn(0.8, second); // :94
}

__global__ void __launch_bounds__(K5_THREADS) k5_count_kernel(const K5Pair *__restrict__ pairs, uint32_t n_pairs,
                                                               unsigned long long *__restrict__ offsets,
                                                               uint32_t *__restrict__ ticket)
{
    __shared__ uint32_t warp_sum[K5_THREADS / 32];
    __shared__ uint32_t last;
    __shared__ unsigned long long carry;
    const uint32_t tid = threadIdx.x;
    const K5Pair P = pairs[blockIdx.x];
    uint32_t mine = 0;
    for (uint32_t i = tid; i < P.n_q; i += K5_THREADS)
        mine += ratio_test_keeps(P.top[i]) ? 1u : 0u;
    mine = __reduce_add_sync(0xFFFFFFFFu, mine);
    if ((tid & 31) == 0)
        warp_sum[tid >> 5] = mine;
    __syncthreads();
    if (tid == 0)
    {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < K5_THREADS / 32; w++)
            total += warp_sum[w];
        offsets[blockIdx.x + 1] = total;
        __threadfence();
        last = atomicAdd(ticket, 1u) == n_pairs - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!last)
        return;
    // the last CTA turns counts into offsets: offsets[p] = sum of counts[0 .. p), offsets[n_pairs] = total
    __threadfence();
    if (tid == 0)
    {
        carry = 0;
        offsets[0] = 0;
    }
    __syncthreads();
    for (uint32_t base = 0; base < n_pairs; base += K5_THREADS)
    {
        const uint32_t p = base + tid;
        unsigned long long v = p < n_pairs ? __ldcg(&offsets[p + 1]) : 0ull;
        // inclusive scan of the tile: warp shuffles, then the warps' totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if ((tid & 31) >= (uint32_t)d)
                v += up;
        }
        __shared__ unsigned long long wtot[K5_THREADS / 32];
        if ((tid & 31) == 31)
            wtot[tid >> 5] = v;
        __syncthreads();
        unsigned long long before = carry;
        for (uint32_t w = 0; w < (tid >> 5); w++)
            before += wtot[w];
        if (p < n_pairs)
            offsets[p + 1] = before + v;
        __syncthreads();
        if (tid == K5_THREADS - 1)
            carry = before + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(K5_THREADS) k5_compact_kernel(const K5Pair *__restrict__ pairs,
                                                                 const unsigned long long *__restrict__ offsets,
                                                                 ocb_match *__restrict__ out)
{
    __shared__ uint32_t warp_cnt[K5_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const K5Pair P = pairs[blockIdx.x];
    ocb_match *__restrict__ dst = out + offsets[blockIdx.x];
    uint32_t running = 0; // survivors of this pair written so far (same value in every thread)
    for (uint32_t base = 0; base < P.n_q; base += K5_THREADS)
    {
        const uint32_t i = base + tid;
        ocb_top2 r;
        r.best_k = 0, r.best_d = OCB_DIST_INF, r.second_d = OCB_DIST_INF;
        if (i < P.n_q)
            r = P.top[i];
        const bool keep = i < P.n_q && ratio_test_keeps(r);
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0)
            warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        uint32_t before = running, total = 0;
#pragma unroll
        for (int w = 0; w < K5_THREADS / 32; w++)
        {
            const uint32_t c = warp_cnt[w];
            before += (uint32_t)w < warp ? c : 0u;
            total += c;
        }
        if (keep)
        {
            ocb_match m;
            m.query_k = i, m.best_k = r.best_k, m.best_d = r.best_d;
            dst[before + __popc(ballot & ((1u << lane) - 1u))] = m;
        }
        running += total;
        __syncthreads();
    }
}

int k5_ratio_compact(const K5Pair *d_pairs, size_t n_pairs, unsigned long long *d_offsets, uint32_t *d_ticket,
                     ocb_match *d_out, cudaStream_t stream)
{
    if (n_pairs == 0)
        return 0;
    OCB_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t), stream));
    k5_count_kernel<<<(unsigned)n_pairs, K5_THREADS, 0, stream>>>(d_pairs, (uint32_t)n_pairs, d_offsets, d_ticket);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    k5_compact_kernel<<<(unsigned)n_pairs, K5_THREADS, 0, stream>>>(d_pairs, d_offsets, d_out);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// K7 -- the reference's two std::sort calls on the device: the match list by distance, descending
// (src/match/match_features.cpp:100-101), and the PROSAC pool by quality, ascending (src/model_inliers/ransac.cpp:83-90).
// Both are unstable sorts whose tie order matters downstream, so libstdc++'s algorithm is replayed step for step
// (std_sort_replay.cuh) on (key, position) words: one warp per pair, the words in shared memory, lane 0 running the
// sequential algorithm while the other lanes load, permute and store. Latency bound (a dependent shared-memory access
// per step, ~1 ms per pair) and off the critical path: the warps of a submission's pairs run next to the K1 CTAs of
// the next submission.
// ----------------------------------------------------------------------------------------------------------
constexpr int K7_THREADS = 32;

__global__ void __launch_bounds__(K7_THREADS)
    k7_sort_kernel(const unsigned long long *__restrict__ offsets, const ocb_match *__restrict__ in,
                   ocb_match *__restrict__ out, uint32_t *__restrict__ quality_order,
                   unsigned long long *__restrict__ scratch, uint32_t smem_cap)
{
    extern __shared__ unsigned long long k7_words[];
    const uint32_t lane = threadIdx.x;
    const unsigned long long lo = offsets[blockIdx.x];
    const uint32_t n = (uint32_t)(offsets[blockIdx.x + 1] - lo);
    if (n == 0)
        return;
    uint64_t *v = reinterpret_cast<uint64_t *>(n <= smem_cap ? k7_words : scratch + lo); // long lists: global scratch
    for (uint32_t i = lane; i < n; i += K7_THREADS)
        v[i] = ((uint64_t)in[lo + i].best_d << 32) | i;
    __syncwarp();
    if (lane == 0)
        sort_replay::std_sort(v, (long)n, sort_replay::KeyOrder<true>());
    __syncwarp();
    for (uint32_t i = lane; i < n; i += K7_THREADS)
        out[lo + i] = in[lo + (uint32_t)v[i]];
    if (!quality_order)
        return;
    // correspondence i belongs to sorted match i and carries quality = its distance
    for (uint32_t i = lane; i < n; i += K7_THREADS)
        v[i] = (v[i] & 0xFFFFFFFF00000000ull) | i;
    __syncwarp();
    if (lane == 0)
        sort_replay::std_sort(v, (long)n, sort_replay::KeyOrder<false>());
    __syncwarp();
    for (uint32_t i = lane; i < n; i += K7_THREADS)
        quality_order[lo + i] = (uint32_t)v[i];
}

int k7_sort(const unsigned long long *d_offsets, size_t n_pairs, uint32_t max_rows_per_pair, const ocb_match *d_in,
            ocb_match *d_out, uint32_t *d_quality_order, unsigned long long *d_scratch, cudaStream_t stream)
{
    if (n_pairs == 0)
        return 0;
    // the words of a pair live in shared memory when they fit (up to 96 KB: two such warps per SM next to K1's CTAs)
    const uint32_t smem_cap = std::min<uint32_t>(std::max<uint32_t>(max_rows_per_pair, 1u), 96u * 1024u / 8u);
    const size_t smem = (size_t)smem_cap * sizeof(unsigned long long);
    OCB_CUDA(cudaFuncSetAttribute(k7_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    k7_sort_kernel<<<(unsigned)n_pairs, K7_THREADS, smem, stream>>>(d_offsets, d_in, d_out, d_quality_order, d_scratch,
                                                                   smem_cap);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// K6
// ----------------------------------------------------------------------------------------------------------
constexpr int K6_THREADS = 128;

// std::max / std::min semantics (the first argument wins ties and NaNs), not fmax / fmin
__device__ __forceinline__ double max_std(double a, double b)
{
    return (a < b) ? b : a;
}
__device__ __forceinline__ double min_std(double a, double b)
{
    return (b < a) ? b : a;
}

// distortProjectedRay (include/opencalibration/distort/distort_keypoints.hpp:27-43) and its 2x2 Jacobian, in the
// operation order of host/distort_keypoints.cpp: distort_ray
template <bool JAC>
__device__ __forceinline__ void distort_ray_dev(const ocb_camera &cm, double x, double y, double *out, double *J)
{
    const double k0 = cm.radial_distortion[0], k1 = cm.radial_distortion[1], k2 = cm.radial_distortion[2];
    const double p0 = cm.tangential_distortion[0], p1 = cm.tangential_distortion[1];
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    const double r4 = __dmul_rn(r2, r2), r6 = __dmul_rn(r4, r2);
    const double radial =
        __dadd_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(k0, r2), __dmul_rn(k1, r4)), __dmul_rn(k2, r6)));
    const double xy = __dmul_rn(x, y);
    const double two_xy = __dmul_rn(2.0, xy);
    out[0] = __dadd_rn(__dadd_rn(__dmul_rn(radial, x), __dmul_rn(two_xy, p0)),
                       __dmul_rn(p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
    out[1] = __dadd_rn(__dadd_rn(__dmul_rn(radial, y), __dmul_rn(two_xy, p1)),
                       __dmul_rn(p0, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))));
    if (JAC)
    {
        // dr = k0 + 2 k1 r2 + 3 k2 r4
        const double dr = __dadd_rn(__dadd_rn(k0, __dmul_rn(__dmul_rn(2.0, k1), r2)), __dmul_rn(__dmul_rn(3.0, k2), r4));
        const double drx = __dmul_rn(__dmul_rn(dr, 2.0), x), dry = __dmul_rn(__dmul_rn(dr, 2.0), y);
        const double two_x = __dmul_rn(2.0, x), two_y = __dmul_rn(2.0, y);
        J[0] = __dadd_rn(__dadd_rn(__dadd_rn(radial, __dmul_rn(x, drx)), __dmul_rn(two_y, p0)),
                         __dmul_rn(p1, __dadd_rn(two_x, __dmul_rn(4.0, x))));
        J[1] = __dadd_rn(__dadd_rn(__dmul_rn(x, dry), __dmul_rn(two_x, p0)), __dmul_rn(p1, two_y));
        J[2] = __dadd_rn(__dadd_rn(__dmul_rn(y, drx), __dmul_rn(two_y, p1)), __dmul_rn(p0, two_x));
        J[3] = __dadd_rn(__dadd_rn(__dadd_rn(radial, __dmul_rn(y, dry)), __dmul_rn(two_x, p1)),
                         __dmul_rn(p0, __dadd_rn(two_y, __dmul_rn(4.0, y))));
    }
}

// ceres::TinySolver<..., 2, 2>::Solve as restated in host/distort_keypoints.cpp: undistort_lm
struct LmState
{
    double x[2], f[2], J[4], scale[2], jtj[4], g[2], cost;
};
__device__ __noinline__ double lm_update(const ocb_camera &cm, const double *target, LmState &s)
{
    double dist[2], Jd[4];
    distort_ray_dev<true>(cm, s.x[0], s.x[1], dist, Jd);
    s.f[0] = __dsub_rn(target[0], dist[0]), s.f[1] = __dsub_rn(target[1], dist[1]);
#pragma unroll
    for (int i = 0; i < 4; i++)
        s.J[i] = -Jd[i];
#pragma unroll
    for (int c = 0; c < 2; c++)
    {
        const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(s.J[c], s.J[c]), __dmul_rn(s.J[2 + c], s.J[2 + c])));
        s.scale[c] = __ddiv_rn(1.0, __dadd_rn(1.0, nrm));
        s.J[c] = __dmul_rn(s.J[c], s.scale[c]), s.J[2 + c] = __dmul_rn(s.J[2 + c], s.scale[c]);
    }
    s.jtj[0] = __dadd_rn(__dmul_rn(s.J[0], s.J[0]), __dmul_rn(s.J[2], s.J[2]));
    s.jtj[1] = s.jtj[2] = __dadd_rn(__dmul_rn(s.J[0], s.J[1]), __dmul_rn(s.J[2], s.J[3]));
    s.jtj[3] = __dadd_rn(__dmul_rn(s.J[1], s.J[1]), __dmul_rn(s.J[3], s.J[3]));
    s.g[0] = -__dadd_rn(__dmul_rn(s.J[0], s.f[0]), __dmul_rn(s.J[2], s.f[1]));
    s.g[1] = -__dadd_rn(__dmul_rn(s.J[1], s.f[0]), __dmul_rn(s.J[3], s.f[1]));
    s.cost = __dmul_rn(0.5, __dadd_rn(__dmul_rn(s.f[0], s.f[0]), __dmul_rn(s.f[1], s.f[1])));
    return max_std(fabs(s.g[0]), fabs(s.g[1]));
}

__device__ __noinline__ void undistort_lm_dev(const ocb_camera &cm, const double *target, double parameter_tolerance,
                                              double *x_io)
{
    const double gradient_tolerance = __dmul_rn(parameter_tolerance, 1e-2), cost_threshold = 1e-16;
    LmState s;
    s.x[0] = x_io[0], s.x[1] = x_io[1];
    if (lm_update(cm, target, s) < gradient_tolerance || s.cost < cost_threshold)
        return;
    double u = 1.0 / 1e4, v = 2.0;
    for (int it = 1; it < 10; it++)
    {
        const double a = __dadd_rn(s.jtj[0], __dmul_rn(u, min_std(max_std(s.jtj[0], 1e-6), 1e32))), b = s.jtj[1];
        const double c = __dadd_rn(s.jtj[3], __dmul_rn(u, min_std(max_std(s.jtj[3], 1e-6), 1e32)));
        const double det = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, b));
        const double s0 = __ddiv_rn(__dsub_rn(__dmul_rn(c, s.g[0]), __dmul_rn(b, s.g[1])), det);
        const double s1 = __ddiv_rn(__dsub_rn(__dmul_rn(a, s.g[1]), __dmul_rn(b, s.g[0])), det);
        const double dx0 = __dmul_rn(s.scale[0], s0), dx1 = __dmul_rn(s.scale[1], s1);
        const double xnorm = __dsqrt_rn(__dadd_rn(__dmul_rn(s.x[0], s.x[0]), __dmul_rn(s.x[1], s.x[1])));
        if (__dsqrt_rn(__dadd_rn(__dmul_rn(dx0, dx0), __dmul_rn(dx1, dx1))) <
            __dmul_rn(parameter_tolerance, __dadd_rn(xnorm, parameter_tolerance)))
            break;
        const double xn0 = __dadd_rn(s.x[0], dx0), xn1 = __dadd_rn(s.x[1], dx1);
        double dist[2];
        distort_ray_dev<false>(cm, xn0, xn1, dist, nullptr);
        const double fn0 = __dsub_rn(target[0], dist[0]), fn1 = __dsub_rn(target[1], dist[1]);
        const double cost_change =
            __dsub_rn(__dmul_rn(2.0, s.cost), __dadd_rn(__dmul_rn(fn0, fn0), __dmul_rn(fn1, fn1)));
        const double mc0 = __dmul_rn(
            s0, __dsub_rn(__dmul_rn(2.0, s.g[0]), __dadd_rn(__dmul_rn(s.jtj[0], s0), __dmul_rn(s.jtj[1], s1))));
        const double mc1 = __dmul_rn(
            s1, __dsub_rn(__dmul_rn(2.0, s.g[1]), __dadd_rn(__dmul_rn(s.jtj[2], s0), __dmul_rn(s.jtj[3], s1))));
        const double rho = __ddiv_rn(cost_change, __dadd_rn(mc0, mc1));
        if (rho > 0)
        {
            s.x[0] = xn0, s.x[1] = xn1;
            if (lm_update(cm, target, s) < gradient_tolerance || s.cost < cost_threshold)
                break;
            const double tmp = __dsub_rn(__dmul_rn(2.0, rho), 1.0);
            u = __dmul_rn(u, max_std(1.0 / 3.0, __dsub_rn(1.0, __dmul_rn(__dmul_rn(tmp, tmp), tmp))));
            v = 2.0;
            continue;
        }
        u = __dmul_rn(u, v);
        v = __dmul_rn(v, 2.0);
    }
    x_io[0] = s.x[0], x_io[1] = s.x[1];
}

// image_to_3d (src/distort/distort_keypoints.cpp:62-103) as restated in host/distort_keypoints.cpp
__device__ __forceinline__ void image_to_3d_dev(const ocb_camera &cm, double px, double py, double *ray)
{
    const double f = cm.focal_length_pixels;
    const double unprojected[2] = {__ddiv_rn(__dsub_rn(px, cm.principal_point[0]), f),
                                   __ddiv_rn(__dsub_rn(py, cm.principal_point[1]), f)};
    double und[2] = {unprojected[0], unprojected[1]};
    if (cm.radial_distortion[0] != 0 || cm.radial_distortion[1] != 0 || cm.radial_distortion[2] != 0 ||
        cm.tangential_distortion[0] != 0 || cm.tangential_distortion[1] != 0)
    {
        const double pp_norm = __dsqrt_rn(__dadd_rn(__dmul_rn(cm.principal_point[0], cm.principal_point[0]),
                                                    __dmul_rn(cm.principal_point[1], cm.principal_point[1])));
        undistort_lm_dev(cm, unprojected, __ddiv_rn(1e-2, __dadd_rn(pp_norm, f)), und);
    }
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    ray[0] = ray[1] = ray[2] = nan; // ProjectionType::UNKNOWN leaves the ray unset (:93-101); the mirror's is NaN
    if (cm.projection_planar)
    {
        const double z = __dadd_rn(__dadd_rn(__dmul_rn(und[0], und[0]), __dmul_rn(und[1], und[1])), 1.0);
        if (z > 0) // Eigen's normalized()
        {
            const double n = __dsqrt_rn(z);
            ray[0] = __ddiv_rn(und[0], n), ray[1] = __ddiv_rn(und[1], n), ray[2] = __ddiv_rn(1.0, n);
        }
        else
            ray[0] = und[0], ray[1] = und[1], ray[2] = 1.0;
    }
}

__global__ void __launch_bounds__(K6_THREADS) k6_rays_kernel(const K6Set *__restrict__ sets, uint32_t n_sets)
{
    // CTA -> set by binary search over cta_begin; thread -> (match, side)
    uint32_t lo = 0, hi = n_sets - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (sets[mid].cta_begin <= blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const K6Set &S = sets[lo];
    const uint32_t local = (blockIdx.x - S.cta_begin) * (K6_THREADS / 2) + (threadIdx.x >> 1);
    const uint32_t side = threadIdx.x & 1;
    if (local >= S.n)
        return;
    const ocb_match m = S.matches[local];
    const double2 px = side ? S.xy2[m.best_k] : S.xy1[m.query_k];
    double ray[3];
    image_to_3d_dev(side ? S.cam2 : S.cam1, px.x, px.y, ray);
    double *row = S.c7 + (size_t)local * 7 + side * 3;
    row[0] = ray[0], row[1] = ray[1], row[2] = ray[2];
    if (side == 0) // correspondence::quality = feature_match::distance (:61) = count * (1.0 / 486)
        S.c7[(size_t)local * 7 + 6] = __dmul_rn((double)m.best_d, 1.0 / OCB_DESCRIPTOR_BITS);
    else if (S.order_src)
        S.order_dst[local] = S.order_src[local];
}

uint32_t k6_set_ctas(uint32_t n)
{
    return (n + K6_THREADS / 2 - 1) / (K6_THREADS / 2);
}

int k6_rays(const K6Set *d_sets, size_t n_sets, uint32_t total_ctas, cudaStream_t stream)
{
    if (n_sets == 0 || total_ctas == 0)
        return 0;
    k6_rays_kernel<<<total_ctas, K6_THREADS, 0, stream>>>(d_sets, (uint32_t)n_sets);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(K6_THREADS) k6_points_kernel(const double2 *__restrict__ xy, uint32_t n,
                                                                const ocb_camera cam, double *__restrict__ rays)
{
    const uint32_t i = blockIdx.x * K6_THREADS + threadIdx.x;
    if (i >= n)
        return;
    const double2 p = xy[i];
    double ray[3];
    image_to_3d_dev(cam, p.x, p.y, ray);
    rays[(size_t)i * 3 + 0] = ray[0], rays[(size_t)i * 3 + 1] = ray[1], rays[(size_t)i * 3 + 2] = ray[2];
}

int k6_points(const double *d_xy, size_t n, const ocb_camera &cam, double *d_rays, cudaStream_t stream)
{
    if (n == 0)
        return 0;
    k6_points_kernel<<<(unsigned)((n + K6_THREADS - 1) / K6_THREADS), K6_THREADS, 0, stream>>>(
        reinterpret_cast<const double2 *>(d_xy), (uint32_t)n, cam, d_rays);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace ocb
