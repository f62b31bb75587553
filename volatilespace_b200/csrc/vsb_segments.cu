// vsb_segments.cu -- phys_vessel.Physics.curve_segments (phys_vessel.py:828-1075) for the
// visible vessels, on the device-resident vessel state (sm_100a).
//
// For every listed vessel the reference cuts its moved orbit curve (P points) into a "light"
// part -- from the vessel to the next thing that happens on the orbit (body impact, COI entry
// or COI leave, whichever comes first in orbit direction) -- and a "dark" part that ends where
// the vessel leaves the COI.  Each part is concat_wrap (phys_vessel.py:35-79): an end point, a
// run of curve rows (wrapping through the end of the array when the range is reversed), the
// other end point, NaN padding; rows are compacted when the source curve holds NaNs.
//
// One CTA per vessel: thread 0 turns the orbit points into a plan (which rows, which end
// points; index arithmetic with rint = Python's round-half-even), all threads then copy rows.
// The curves never leave the device except for the (count, P, 2) results.
// Compiled with -fmad=false.
#include "vsb_common.cuh"
#include "vsb_orbit_math.cuh"

namespace {

struct seg_plan {
    int mode;            // 0: copy the whole curve, 1: concat_wrap, 2: not written (NaN rows)
    int r0, r1;
    double first[2], last[2];
};

// numpy slice bound: negative counts from the end, clamped to [0, P]
__device__ __forceinline__ int slice_bound(int i, int P)
{
    if (i < 0) {
        i += P;
        return i < 0 ? 0 : i;
    }
    return i > P ? P : i;
}

__device__ __forceinline__ int curve_index(double ea, int P, bool ell)
{
    // phys_vessel.py:866-867 / :877-878
    if (ell) return (int)rint(ea * P / (2 * PI_D));
    return P - (int)rint((ea + PI_D) * P / (2 * PI_D));
}

__device__ void plan_wrap(seg_plan &s, const double *a, int r0, int r1, const double *b)
{
    s.mode = 1;
    s.r0 = r0;
    s.r1 = r1;
    s.first[0] = a[0]; s.first[1] = a[1];
    s.last[0] = b[0]; s.last[1] = b[1];
}

// hyperbola: the curve array runs from ea = +pi down to -pi (phys_vessel.py:875-893, 929-946,
// 992-1003); `clamp` is the COI-leave form
__device__ void plan_hyp_light(seg_plan &s, double ea, int ev, int en, bool ccw, const double *here,
                               const double *there, bool clamp)
{
    if (ccw) {
        if (ev == en) ev += 1;
        int hi = ev - 1;
        if (clamp && hi < 0) hi = 0;
        if (clamp || ea < 0) plan_wrap(s, there, en, hi, here);
        else plan_wrap(s, here, en, hi, there);
        return;
    }
    if (ev == en) ev -= 1;
    if (clamp || !(ea < 0)) plan_wrap(s, here, ev + 1, en, there);
    else plan_wrap(s, there, ev + 1, en, here);
}

// writes one part (P rows) of one vessel
__device__ void emit_part(const seg_plan &s, const double2 *__restrict__ src, int P, bool src_has_nan,
                          double2 *__restrict__ out, int *sh_scan)
{
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (s.mode != 1) {
        for (int j = threadIdx.x; j < P; j += blockDim.x)
            out[j] = s.mode == 0 ? src[j] : make_double2(nan, nan);
        return;
    }
    // row j of the un-compacted result
    int cnt, start, n_head, lead;      // lead: 0 = `first` leads, 1 = `last` leads
    if (s.r0 <= s.r1) {
        cnt = min(s.r1 - s.r0, P - 2);
        start = slice_bound(s.r0, P);
        const int len = max(slice_bound(s.r0 + cnt, P) - start, 0);
        n_head = min(len, cnt);   // rows actually available (== cnt for every valid input)
        lead = 0;
    } else {
        start = slice_bound(s.r0, P);
        const int len1 = P - start, len2 = slice_bound(s.r1, P);
        cnt = min(len1 + len2, P - 2);
        n_head = min(cnt, len1);
        lead = 1;
    }
    auto row = [&](int j) -> double2 {
        if (j == 0) return lead ? make_double2(s.last[0], s.last[1]) : make_double2(s.first[0], s.first[1]);
        if (j == 1 + cnt) return lead ? make_double2(s.first[0], s.first[1]) : make_double2(s.last[0], s.last[1]);
        if (j > 1 + cnt) return make_double2(nan, nan);
        const int q = j - 1;
        if (q < n_head) return src[start + q];
        if (lead) return src[q - n_head];          // the part before r1, from the start of the array
        return make_double2(nan, nan);
    };
    if (!src_has_nan) {
        for (int j = threadIdx.x; j < P; j += blockDim.x) out[j] = row(j);
        return;
    }
    // compaction of the rows whose x is not NaN (stable), NaN rows behind them
    int written = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < P; base += blockDim.x) {
        const int j = base + threadIdx.x;
        double2 r = make_double2(nan, nan);
        if (j < P) r = row(j);
        const bool keep = j < P && r.x == r.x;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sh_scan[w] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < nw; ++k) {
            if (k < w) before += sh_scan[k];
            total += sh_scan[k];
        }
        if (keep) out[written + before + __popc(bal & ((1u << lane) - 1))] = r;
        written += total;
        __syncthreads();
    }
    for (int j = written + threadIdx.x; j < P; j += blockDim.x) out[j] = make_double2(nan, nan);
}

__global__ void __launch_bounds__(128)
curve_segments_kernel(const int64_t *__restrict__ idx, int P, const double *__restrict__ a,
                      const double *__restrict__ b, const double *__restrict__ f,
                      const double *__restrict__ ecc, const double *__restrict__ pea,
                      const double *__restrict__ nmot, const double *__restrict__ dr,
                      const double *__restrict__ ma, const double *__restrict__ ea,
                      const double2 *__restrict__ pos, const int64_t *__restrict__ ref,
                      const double2 *__restrict__ ref_pos, const double *__restrict__ impact,
                      const double *__restrict__ leave, const double *__restrict__ enter,
                      const double2 *__restrict__ curves, double2 *__restrict__ light,
                      double2 *__restrict__ dark, double *__restrict__ first_out,
                      int *__restrict__ type_out, double *__restrict__ time_out,
                      double *__restrict__ range_out)
{
    __shared__ seg_plan plan[2];
    __shared__ int sh_scan[4];
    const int slot = blockIdx.x;
    const int64_t v = idx[slot];
    const double2 *src = curves + (size_t)v * P;
    // np.any(np.isnan(array)), phys_vessel.py:73
    int bad = 0;
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        const double2 r = src[j];
        bad |= (r.x != r.x) || (r.y != r.y);
    }
    const bool src_has_nan = __syncthreads_or(bad) != 0;
    if (threadIdx.x == 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const double A = a[v], B = b[v], F = f[v], E = ecc[v], W = pea[v], N = nmot[v], D = dr[v],
                     M = ma[v], EAV = ea[v];
        const bool ell = E < 1, ccw = D > 0;
        const int col = ell ? 0 : 1;
        const double here[2] = {pos[v].x, pos[v].y};
        const double2 rp = ref_pos[ref[v]];
        // what comes next: impact, COI entry, COI leave (phys_vessel.py:855-859)
        const double cand[3] = {impact[2 * v + col], enter[6 * v + 1], leave[2 * v + col]};
        const int nearest = first_intersect(EAV, cand, 3, D);
        seg_plan lp, dp;
        lp.mode = 0;
        dp.mode = 2;
        double first[2] = {nan, nan}, t = 0.0, range[2] = {nan, nan};
        int type = 0;
        const int ev0 = curve_index(EAV, P, ell);
        if (nearest == 0 || nearest == 1) {
            dp.mode = 0;
            const double ea_next = nearest == 0 ? impact[2 * v] : enter[6 * v + 1];
            double there[2];
            ps_orb2xy(A, B, F, E, W, rp.x, rp.y, ea_next, there);
            int ev = ev0;
            const int en = curve_index(ea_next, P, ell);
            if (!ell) {
                plan_hyp_light(lp, EAV, ev, en, ccw, here, there, false);
            } else if (nearest == 0) {
                if (ccw) {
                    if (ev == en) ev -= 1;
                    plan_wrap(lp, here, ev + 1, en, there);
                } else {
                    if (ev == en) ev += 1;
                    plan_wrap(lp, there, en, ev - 1, here);
                }
            } else if (ccw) {
                if (EAV < PI_D) {
                    if (ev == en) ev -= 1;
                    plan_wrap(lp, here, ev + 1, en, there);
                } else if (PI_D < EAV && EAV < PI_D * 1.5) {
                    plan_wrap(lp, here, ev + 1, en, there);
                } else {
                    plan_wrap(lp, there, ev, en, here);
                }
            } else {
                if (!(EAV < PI_D) && ev == en) ev += 1;
                plan_wrap(lp, there, en, ev - 1, here);
            }
            first[0] = there[0]; first[1] = there[1];
            t = orbit_time_to(M, ea_next - E * sin(ea_next), E, D, N);
            type = 1 + nearest;
            if (nearest == 1) {
                range[0] = EAV;
                range[1] = ea_next;
            }
        }
        // does a COI leave lie anywhere ahead?  (`any(next_intersect == 2)`: its entry is not NaN)
        const bool leave_ahead = nearest >= 0 && cand[2] == cand[2];
        if (nearest == 2 || leave_ahead) {
            const double ea_next = leave[2 * v], ea_prev = leave[2 * v + 1];
            double there[2], behind[2];
            ps_orb2xy(A, B, F, E, W, rp.x, rp.y, ea_next, there);
            ps_orb2xy(A, B, F, E, W, rp.x, rp.y, ea_prev, behind);
            int ev = ev0;
            const int en = curve_index(ea_next, P, ell), ep = curve_index(ea_prev, P, ell);
            if (nearest == 2) {
                double ma_next;
                if (ell) {
                    ma_next = ea_next - E * sin(ea_next);
                    if (ccw) {
                        if (EAV < PI_D) {
                            if (ev == en) ev -= 1;
                            plan_wrap(lp, here, ev + 1, en, there);
                        } else {
                            plan_wrap(lp, there, ev, en, here);
                        }
                        plan_wrap(dp, there, ep, en, behind);
                    } else {
                        if (EAV < PI_D) {
                            plan_wrap(lp, here, en, max(ev, 0), there);
                        } else {
                            if (ev == en) ev += 1;
                            plan_wrap(lp, there, en, max(ev - 1, 0), here);
                        }
                        plan_wrap(dp, behind, en, ep, there);
                    }
                } else {
                    ma_next = (E * sinh(ea_next) - ea_next) * -1;
                    plan_hyp_light(lp, EAV, ev, en, ccw, here, there, true);
                    if (ccw) plan_wrap(dp, there, en, ep, behind);
                    else plan_wrap(dp, behind, ep, en, there);
                }
                first[0] = there[0]; first[1] = there[1];
                t = orbit_time_to(M, ma_next, E, D, N);
                type = 3;
            } else if (ell) {
                if (ccw) plan_wrap(dp, there, ep, en, behind);
                else if (EAV < PI_D) plan_wrap(dp, behind, en, ep, there);
                else plan_wrap(dp, there, en, ep, behind);
            } else if (ccw) {
                plan_wrap(dp, there, en, ep, behind);
            } else {
                plan_wrap(dp, behind, ep, en, there);
            }
        }
        plan[0] = lp;
        plan[1] = dp;
        first_out[2 * slot] = first[0];
        first_out[2 * slot + 1] = first[1];
        type_out[slot] = type;
        time_out[slot] = t;
        range_out[2 * slot] = range[0];
        range_out[2 * slot + 1] = range[1];
    }
    __syncthreads();
    emit_part(plan[0], src, P, src_has_nan, light + (size_t)slot * P, sh_scan);
    __syncthreads();
    emit_part(plan[1], src, P, src_has_nan, dark + (size_t)slot * P, sh_scan);
}

}  // namespace

// d_idx: `count` vessel indices on the device; outputs on the device
int vsb_launch_curve_segments(vsb_ctx *c, const vsb_kepler_state &k, const int64_t *d_idx, int64_t count,
                              double *light, double *dark, double *first, int *type, double *time,
                              double *range)
{
    if (count <= 0) return VSB_OK;
    curve_segments_kernel<<<(unsigned)count, 128, 0, c->stream>>>(
        d_idx, (int)k.P, k.a, k.b, k.f, k.ecc, k.pea, k.nmot, k.dr, k.ma, k.ea, (const double2 *)k.pos,
        k.ref, (const double2 *)k.ref_pos, k.ev_impact, k.ev_leave, k.ev_enter,
        (const double2 *)k.curves.p, (double2 *)light, (double2 *)dark, first, type, time, range);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
