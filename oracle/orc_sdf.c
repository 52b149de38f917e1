/* orc_sdf.c -- CPU ORACLE (test infrastructure): procedural SDFs and the per-chunk generators.
 * Restates Runtimes/Helper/GeneratorHelper.h:19-150 and Runtimes/Helper/VoxelMathHelper.h:25-33
 * line by line in fp64 with glm's evaluation order (left-associative, component-wise), compiled
 * with -ffp-contract=off.  See meso_oracle.h for the parity status. */
#include "orc_internal.h"

/* ---- portable sin ------------------------------------------------------------------------
 * libm sin is not bit-reproducible across platforms (the reference itself runs MSVC's), nor on
 * the GPU.  ORC_SIN_PORTABLE is a fixed sequence of fp64 + - * floor only (4-term Cody-Waite
 * reduction by pi/2 in 30-bit pieces, fdlibm-style minimax kernels), restated operation for
 * operation by the CUDA voxeliser.  Valid for |x| < 2^23 * pi/2 (terrain hashes stay < 1e7). */
static const double PIO2_1 = 0x1.921fb54000000p+0;
static const double PIO2_2 = 0x1.10b4611800000p-30;
static const double PIO2_3 = 0x1.313198a000000p-61;
static const double PIO2_4 = 0x1.701b839a25205p-92;
static const double TWO_OVER_PI = 0x1.45f306dc9c883p-1;
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                    S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                    S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                    C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                    C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

double orc_sin_portable(double x) {
  double k = floor(x * TWO_OVER_PI + 0.5);
  double r = x - k * PIO2_1;
  r = r - k * PIO2_2;
  r = r - k * PIO2_3;
  r = r - k * PIO2_4;
  /* quadrant = k mod 4, computed in fp64 (k is an exact integer) */
  double q = k - 4.0 * floor(k * 0.25);
  double z = r * r;
  double s, c;
  {
    double v = z * r;
    double p = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    s = r + v * (S1 + z * p);
  }
  {
    double p = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    c = 1.0 - (0.5 * z - z * p);
  }
  if (q == 0.0) return s;
  if (q == 1.0) return c;
  if (q == 2.0) return -s;
  return -c;
}

static inline double orc_sin(double x, int sin_mode) { return sin_mode == ORC_SIN_LIBM ? sin(x) : orc_sin_portable(x); }

/* VoxelMathHelper.h:25-33: Hash(p) = Fract(sin(dot(p,(127.1,311.7,74.7))) * 43758.5453123),
 * Fract(v) = v - floor(v); glm dot(dvec3) = (x*x' + y*y') + z*z'. */
double orc_hash3(double x, double y, double z, int sin_mode) {
  double d = (x * 127.1 + y * 311.7) + z * 74.7;
  double v = orc_sin(d, sin_mode) * 43758.5453123;
  return v - floor(v);
}

/* GeneratorHelper.h:19-53.  out4 = (grad.x, grad.y, grad.z, value) -- the ".yzwx" swizzle. */
void orc_noised(const double x[3], int sin_mode, double out4[4]) {
  double p[3], w[3], u[3], du[3];
  for (int i = 0; i < 3; i++) {
    p[i] = floor(x[i]);
    w[i] = x[i] - floor(x[i]);                                   /* glm::fract */
    u[i] = ((w[i] * w[i]) * w[i]) * ((w[i] * ((w[i] * 6.0) - 15.0)) + 10.0);
    du[i] = ((30.0 * w[i]) * w[i]) * ((w[i] * (w[i] - 2.0)) + 1.0);
  }
  double a = orc_hash3(p[0] + 0, p[1] + 0, p[2] + 0, sin_mode);
  double b = orc_hash3(p[0] + 1, p[1] + 0, p[2] + 0, sin_mode);
  double c = orc_hash3(p[0] + 0, p[1] + 1, p[2] + 0, sin_mode);
  double d = orc_hash3(p[0] + 1, p[1] + 1, p[2] + 0, sin_mode);
  double e = orc_hash3(p[0] + 0, p[1] + 0, p[2] + 1, sin_mode);
  double f = orc_hash3(p[0] + 1, p[1] + 0, p[2] + 1, sin_mode);
  double g = orc_hash3(p[0] + 0, p[1] + 1, p[2] + 1, sin_mode);
  double h = orc_hash3(p[0] + 1, p[1] + 1, p[2] + 1, sin_mode);
  double k0 = a;
  double k1 = b - a;
  double k2 = c - a;
  double k3 = e - a;
  double k4 = a - b - c + d;
  double k5 = a - c - e + g;
  double k6 = a - b - e + f;
  double k7 = -a + b + c - d + e - f - g + h;
  double val = -1.0 + 2.0 * (k0 + k1 * u[0] + k2 * u[1] + k3 * u[2] + k4 * u[0] * u[1] + k5 * u[1] * u[2] +
                             k6 * u[2] * u[0] + k7 * u[0] * u[1] * u[2]);
  double gx = (2.0 * du[0]) * (k1 + k4 * u[1] + k6 * u[2] + k7 * u[1] * u[2]);
  double gy = (2.0 * du[1]) * (k2 + k5 * u[2] + k4 * u[0] + k7 * u[2] * u[0]);
  double gz = (2.0 * du[2]) * (k3 + k6 * u[0] + k5 * u[1] + k7 * u[0] * u[1]);
  out4[0] = gx; out4[1] = gy; out4[2] = gz; out4[3] = val;
}

/* GeneratorHelper.h:56-87 */
double orc_displacement(const double p_in[3], int sin_mode) {
  double p[3] = {p_in[0], p_in[1], p_in[2]};
  double mgn = 0.5, d = 0.0, s = 1.0;
  double rnd[4], q[3];
  for (int i = 0; i < 5; i++) {
    q[0] = p[0] + 10.0; q[1] = p[1] + 10.0; q[2] = p[2] + 10.0;
    orc_noised(q, sin_mode, rnd);
    d += rnd[3] * mgn;
    for (int k = 0; k < 3; k++) { p[k] *= 2.0; p[k] += (rnd[k] * 0.2) * s; }
    if (i == 2) s *= -1.0;
    mgn *= 0.5;
  }
  double sc = pow(2.0, 5);
  p[0] = p_in[0] * sc; p[1] = p_in[1] * sc; p[2] = p_in[2] * sc;
  for (int i = 0; i < 4; i++) {
    orc_noised(p, sin_mode, rnd);
    d += rnd[3] * mgn;
    for (int k = 0; k < 3; k++) p[k] *= 2.0;
    mgn *= 0.5;
  }
  return d;
}

/* GeneratorHelper.h:101-105 (terrain) and :131-135 (sphere; params = centre xyz, radius;
 * the reference hard-codes (100,0,0), 50).  glm length(dvec3) = sqrt((x*x + y*y) + z*z). */
double orc_sdf(int kind, const double params[4], int sin_mode, double x, double y, double z) {
  if (kind == ORC_SDF_SPHERE) {
    double dx = x - params[0], dy = y - params[1], dz = z - params[2];
    return sqrt((dx * dx + dy * dy) + dz * dz) - params[3];
  }
  double q[3] = {x * .1, y * .1, z * .1};
  return (y * .5 + orc_displacement(q, sin_mode) * 10.3) * .4;
}

/* GeneratorHelper.h:90-150: X outer, Z inner; sample at the block MIN corner. */
int orc_generate_chunk(int kind, const double params[4], int sin_mode, const int32_t loc[3], float block_size,
                       int chunk_res, uint8_t* out_xyz) {
  int n = 0;
  for (uint32_t X = 0; X < (uint32_t)chunk_res; X++)
    for (uint32_t Y = 0; Y < (uint32_t)chunk_res; Y++)
      for (uint32_t Z = 0; Z < (uint32_t)chunk_res; Z++) {
        double cs[3], bc[3];
        for (int k = 0; k < 3; k++) cs[k] = (double)loc[k] * (double)block_size * (double)chunk_res;
        bc[0] = cs[0] + (double)X * (double)block_size;
        bc[1] = cs[1] + (double)Y * (double)block_size;
        bc[2] = cs[2] + (double)Z * (double)block_size;
        double d = orc_sdf(kind, params, sin_mode, bc[0], bc[1], bc[2]);
        if (d < 0.0) {
          out_xyz[3 * n + 0] = (uint8_t)X; out_xyz[3 * n + 1] = (uint8_t)Y; out_xyz[3 * n + 2] = (uint8_t)Z;
          n++;
        }
      }
  return n;
}

/* VoxelMathHelper.h:49-71 */
void orc_fibonacci_sphere(uint32_t samples, int normalize, double* out) {
  const double Phi = M_PI * (sqrt(5.0) - 1.0);
  for (int i = 0; i < (int)samples; ++i) {
    double Y = 1 - (i / (double)(samples - 1)) * 2;
    double Radius = sqrt(1 - Y * Y);
    double Theta = Phi * i;
    double X = cos(Theta) * Radius;
    double Z = sin(Theta) * Radius;
    if (normalize) {
      double inv = 1.0 / sqrt((X * X + Y * Y) + Z * Z);  /* glm normalize = v * inversesqrt(dot(v,v)) */
      X *= inv; Y *= inv; Z *= inv;
    }
    out[3 * i] = X; out[3 * i + 1] = Y; out[3 * i + 2] = Z;
  }
}
