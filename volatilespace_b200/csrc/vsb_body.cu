// vsb_body.cu -- SURVEY 8f row N4: the per-frame O(N) body passes of the editor (sm_100a):
//   Physics.body        phys_editor.py:577-599  (surface gravity, atmosphere height, thermal
//                                                luminosity / temperature, Schwarzschild radius)
//   Physics.temp_color  phys_editor.py:602-651  (colour from temperature)
//   Physics.classify    phys_editor.py:654-670  (body type, stellar class)
// One thread per body, structure-of-arrays, one pass.  `rad` is an input (host np.cbrt, it
// feeds the bit-exact collision predicate); physical constants are passed in by the caller so
// they carry the reference's own roundings (phys_shared.py:13-17).  -fmad=false.
#include "vsb_common.cuh"

namespace {

struct body_consts {
    double gc_sim, rad_mult, mass_thermal_mult, min_mass;
    double gc_real, c_light, ls, ms;            // phys_shared.py:13-17
    double sqrt2, pi_q, sigma_q, four_pi;       // np.sqrt(2), np.pi**(1/4), sigma**(1/4), 4*np.pi
};

__global__ void __launch_bounds__(256)
body_thermal_kernel(int64_t n, body_consts K, const double *__restrict__ mass,
                    const double *__restrict__ den, const double *__restrict__ rad,
                    const double *__restrict__ atm_pres0, const double *__restrict__ atm_scale_h,
                    const double *__restrict__ atm_den0, double *__restrict__ atm_h,
                    const long long *__restrict__ base_color, double *__restrict__ surf_grav,
                    double *__restrict__ luminosity, double *__restrict__ temp,
                    double *__restrict__ rad_sc, long long *__restrict__ color,
                    double *__restrict__ types, int *__restrict__ stellar)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = mass[i], d = den[i], r = rad[i];
    // body(), :583-599
    surf_grav[i] = K.gc_sim * m / (r * r);
    double ah = atm_h[i];
    if (atm_pres0[i] != 0 && atm_scale_h[i] != 0 && atm_den0[i] != 0)
        ah = -atm_scale_h[i] * log(0.001 / atm_den0[i]) * K.rad_mult;
    const double thermal_mass = m * K.mass_thermal_mult;
    const double thermal_volume = thermal_mass / d;
    const double thermal_radius = cbrt(3 * thermal_volume / K.four_pi);
    const double lum = K.ls * pow(thermal_mass / K.ms, 3.5);
    const double T = pow(lum, 0.25) / (sqrt(thermal_radius) * K.sqrt2 * K.pi_q * K.sigma_q);
    double rs = 2 * thermal_mass * K.gc_real / (K.c_light * K.c_light);
    rs *= r / thermal_radius;
    if (rs > r) ah = 0.0;
    atm_h[i] = ah;
    luminosity[i] = lum;
    temp[i] = T;
    rad_sc[i] = rs;

    // temp_color(), :602-651 -- every assignment into the int colour array truncates
    const long long b0 = base_color[3 * i], b1 = base_color[3 * i + 1], b2 = base_color[3 * i + 2];
    long long c0 = b0, c1 = b1, c2 = b2;
    if (T > 1000 && T <= 2300) {
        const double fg = 1 - 1 / (1300 / (T - 1000));
        if (fg != -1) {
            c0 = (long long)(b0 * fg + 255 * (1 - fg));
            c1 = (long long)(b1 * fg + 184 * (1 - fg));
            c2 = (long long)(b2 * fg + 111 * (1 - fg));
        }
    }
    if (T > 2300) c0 = (long long)(130 + 125 / pow(1 + pow(T / 5922, 55.525), 0.065));
    if (T > 2300 && T <= 6000) c1 = (long long)(257 - 73 / (1 + pow(T / 4742, 6.687)));
    if (T > 6000) c1 = (long long)(170 + 165766330 / (1 + pow(T / 24, 2.651)));
    if (T > 2300) c2 = (long long)(283 - 176 / (1 + pow(T / 4493, 5.792)));
    if (rs > r) c0 = c1 = c2 = 0;   // black hole
    color[3 * i] = c0 < 0 ? 0 : (c0 > 255 ? 255 : c0);
    color[3 * i + 1] = c1 < 0 ? 0 : (c1 > 255 ? 255 : c1);
    color[3 * i + 2] = c2 < 0 ? 0 : (c2 > 255 ? 255 : c2);

    // classify(), :654-670
    double ty = 0;
    if (m > K.min_mass) ty = 1;
    if (d < 1500) ty = 2;
    if (T > 1000) ty = 3;
    if (rs > r) ty = 4;
    types[i] = ty;
    int sc = ty == 3 ? 1 : 0;   // "M" : "Unknown"
    if (T > 3900) sc = 2;       // K
    if (T > 5300) sc = 3;       // G
    if (T > 6000) sc = 4;       // F
    if (T > 7300) sc = 5;       // A
    if (T > 10000) sc = 6;      // B
    if (T > 33000) sc = 7;      // O
    stellar[i] = sc;
}

}  // namespace

int vsb_launch_body_thermal(vsb_ctx *c, int64_t n, const double *consts12, const double *d_in6,
                            double *d_atm_h, const long long *d_base_color, double *d_out4,
                            long long *d_color, double *d_types, int *d_stellar)
{
    if (n <= 0) return VSB_OK;
    body_consts K;
    memcpy(&K, consts12, sizeof(K));
    body_thermal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        n, K, d_in6, d_in6 + n, d_in6 + 2 * n, d_in6 + 3 * n, d_in6 + 4 * n, d_in6 + 5 * n, d_atm_h,
        d_base_color, d_out4, d_out4 + n, d_out4 + 2 * n, d_out4 + 3 * n, d_color, d_types,
        d_stellar);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
