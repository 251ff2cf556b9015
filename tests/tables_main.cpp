// tests/tables_main.cpp -- host-side checks of the tensor-core operand tables (csrc/b200_tables.cpp:
// b200_build_vseg_mma_tables, b200_build_cnn_mma_tables) against the float weights they were built from.  No GPU.
// usage: tables_main <weights dir>      prints one "name value" line per check; exit code 0 iff all bounds hold
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "b200_internal.h"

static std::vector<float> blob(const std::string &path, size_t n) {
  std::vector<float> v(n);
  FILE *f = fopen(path.c_str(), "rb");
  if (!f || fread(v.data(), 4, n, f) != n) {
    fprintf(stderr, "cannot read %s\n", path.c_str());
    exit(2);
  }
  fclose(f);
  return v;
}
static float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu;
  if (e == 0) return (sign ? -1.0f : 1.0f) * (float)m * 5.9604644775390625e-08f;
  const uint32_t x = sign | ((e + 112u) << 23) | (m << 13);
  float out;
  memcpy(&out, &x, 4);
  return out;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  int bad = 0;
  {  // ---- vseg: W1[u][k] == cu * (((d0 * 128 + d1) * 128 + d2) * 128 + d3) to 2^-27 of the unit's largest weight
    const std::vector<float> w = blob(dir + "/modelm_befe75da.bin", 10403);
    std::vector<int8_t> wq((size_t)4 * 14 * 64 * 16);
    std::vector<VsegUnit> units(64);
    std::vector<float> sd((size_t)256 * 256 * 2);
    b200_build_vseg_mma_tables(w.data(), wq.data(), units.data(), sd.data());
    double worst = 0.0, worst_sum = 0.0;
    int digit_range = 0, pad_nonzero = 0;
    for (int u = 0; u < 64; u++) {
      double smax = 0.0, sum = 0.0;
      for (int k = 0; k < 204 && u < 50; k++) smax = fmax(smax, fabs((double)w[u * 204 + k])), sum += w[u * 204 + k];
      for (int k = 0; k < 224; k++) {
        long long q = 0;
        for (int t = 0; t < 4; t++) {
          const int d = wq[(((size_t)t * 14 + k / 16) * 64 + u) * 16 + k % 16];
          if (d < -64 || d > 64) digit_range++;
          q = q * 128 + d;
        }
        if (u >= 50 || k >= 204) {
          pad_nonzero += q != 0;
          continue;
        }
        worst = fmax(worst, fabs((double)units[u].cu * (double)q - (double)w[u * 204 + k]) / smax);
      }
      if (u < 50) worst_sum = fmax(worst_sum, fabs((double)units[u].sumw - sum));
    }
    printf("vseg_weight_rel_err %.3e\nvseg_digit_out_of_range %d\nvseg_padding_nonzero %d\nvseg_sumw_err %.3e\n", worst, digit_range, pad_nonzero, worst_sum);
    bad += !(worst <= 1.0 / 134217728.0 * 1.01) + (digit_range != 0) + (pad_nonzero != 0) + !(worst_sum < 1e-5);
    // (s, d0): the reference's x = fl(fl(fl(v * k255) * scale) + shift) against (v - mn) * s + d0
    std::vector<float> norm((size_t)256 * 256 * 2);
    b200_build_minmax_norm_table(norm.data());
    const float k255 = 1.0f / 255.0f;
    double worst_x = 0.0;
    for (int mn = 0; mn < 256; mn += 3)
      for (int mx = mn; mx < 256; mx += 5)
        for (int v = mn; v <= mx; v++) {
          const float fs = norm[(mn * 256 + mx) * 2], fb = norm[(mn * 256 + mx) * 2 + 1];
          volatile float a = (float)v * k255;
          volatile float b = a * fs;
          const float x_ref = b + fb;
          const double x = (double)(v - mn) * (double)sd[(mn * 256 + mx) * 2] + (double)sd[(mn * 256 + mx) * 2 + 1];
          // three float roundings of values up to mx / (mx - mn) separate the two
          const double tol = 3.0 * 5.97e-8 * fmax(1.0, (double)mx / fmax(1.0, (double)(mx - mn)));
          if (fabs(x - (double)x_ref) > tol) bad++;
          worst_x = fmax(worst_x, fabs(x - (double)x_ref) / tol);
        }
    printf("vseg_x_vs_reference_in_units_of_tolerance %.3f\n", worst_x);
  }
  {  // ---- digit CNNs
    static const char *names[3] = {"modelc_5c241121.bin", "modelc_01266c1b.bin", "modelc_b00bf70c.bin"};
    std::vector<float> b[3];
    const float *ptrs[3];
    for (int m = 0; m < 3; m++) b[m] = blob(dir + "/" + names[m], 10682), ptrs[m] = b[m].data();
    std::vector<int8_t> convb((size_t)3 * 3 * 2 * 80 * 16);
    std::vector<float> convf(48);
    std::vector<uint16_t> hidb((size_t)3 * 2 * 40 * 32 * 8);
    b200_build_cnn_mma_tables(ptrs, convb.data(), convf.data(), hidb.data());
    double worst_conv = 0.0, worst_hid = 0.0;
    long long worst_sum = 0;
    int misplaced = 0;
    const double k255 = (double)(1.0f / 255.0f);
    for (int m = 0; m < 3; m++) {
      for (int k = 0; k < 8; k++) {
        double smax = 0.0;
        for (int t = 0; t < 9; t++) smax = fmax(smax, fabs((double)b[m][k * 9 + t]));
        for (int pr = 0; pr < 3; pr++)
          for (int pc = 0; pc < 3; pc++) {
            long long abs_sum = 0;
            for (int kk = 0; kk < 32; kk++) {
              const int wy = kk / 5, wx = kk % 5, n = k * 9 + pr * 3 + pc;
              long long q = 0;
              for (int j = 0; j < 3; j++) q = q * 128 + convb[((((size_t)m * 2 + kk / 16) * 240) + 80 * j + n) * 16 + kk % 16];
              const bool tap = kk < 25 && wy >= pr && wy < pr + 3 && wx >= pc && wx < pc + 3;
              if (!tap) {
                misplaced += q != 0;
                continue;
              }
              const double w = b[m][k * 9 + (wy - pr) * 3 + (wx - pc)];
              worst_conv = fmax(worst_conv, fabs((double)convf[m * 8 + k] / k255 * (double)q - w) / smax);
              abs_sum += q < 0 ? -q : q;
            }
            worst_sum = abs_sum > worst_sum ? abs_sum : worst_sum;
          }
      }
      for (int u = 0; u < 32; u++)
        for (int j = 0; j < 320; j++) {
          const int k = j / 40, c = j % 40;
          const size_t at = (((size_t)c * 32) + u) * 8 + k;
          const double got = (double)half_to_float(hidb[((size_t)m * 2) * 10240 + at]) + (double)half_to_float(hidb[((size_t)m * 2 + 1) * 10240 + at]);
          const double w = b[m][80 + u * 320 + j];
          // lo = fp16(w - hi): 2^-22 |w| when normal, half a subnormal step (2.98e-8) otherwise
          worst_hid = fmax(worst_hid, fabs(got - w) / (3.0e-8 + 2.4e-7 * fabs(w)));
        }
    }
    printf("cnn_conv_weight_rel_err %.3e\ncnn_conv_misplaced_taps %d\ncnn_conv_255_sum_abs_q %lld\ncnn_hidden_split_err_in_units_of_bound %.3f\n", worst_conv, misplaced,
           255 * worst_sum, worst_hid);
    bad += !(worst_conv <= 1.0e-6) + (misplaced != 0) + !(255 * worst_sum < 2147483647LL) + !(worst_hid <= 1.0);
  }
  printf("failed_checks %d\n", bad);
  return bad != 0;
}
