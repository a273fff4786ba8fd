// Host build of crime_b200/csrc/gh_group_math.cuh (test infrastructure, not part of the product): runs the grouped
// Taylor pixelisation of accumulate_kernel on the CPU so that tests/test_group_pixelisation_cpu.py can compare every
// accepted answer with the oracle without a GPU.  Built by __graft_entry__.build() with g++ -ffp-contract=off.
#include "../../crime_b200/csrc/gh_group_math.cuh"

extern "C" int gh_group_emulate(double dx, const double *centres, long n, int nside, const float *sub_off_f, float eps_scale,
                                int *kind_out, int *pix_out, unsigned char *ok_out, float *margin_out)
{
  const float fns = (float)nside;
  const int ns = nside, ns4 = 4 * ns;
  const long long npix = 12LL * ns * ns;
  const float hg = (float)(dx * 1.7320508) + 1e-3f;
  for (long i = 0; i < n; ++i) {
    GhGroupExp g;
    const double X = centres[3 * i], Y = centres[3 * i + 1], Z = centres[3 * i + 2];
    const float rc = sqrtf((float)(X * X + Y * Y + Z * Z));
    gh_group_expand(X, Y, Z, hg + 1e-6f * rc, fns, eps_scale, g);
    kind_out[i] = g.kind;
    margin_out[i] = g.kind ? g.e : 0.f;
    if (!g.kind) continue;
    const float hm = 0.5f - g.e;
    for (int c = 0; c < 8; ++c) {
      GhCellExp ce;
      const float px = (float)(((c & 1) - 0.5) * dx), py = (float)((((c >> 1) & 1) - 0.5) * dx), pz = (float)((((c >> 2) & 1) - 0.5) * dx);
      gh_group_recentre(g, px, py, pz, ce);
      for (int s = 0; s < 10; ++s) {
        const float ox = sub_off_f[s], oy = sub_off_f[10 + s], oz = sub_off_f[20 + s];
        const float U = gh_cell_U(g, ce, ox, oy), V = gh_cell_V(g, ce, ox, oy, oz);
        int pix;
        bool ok;
        if (g.kind == GH_GRP_EQ) {
          const int pix0 = 2 * ns * (ns - 1) + ns * ns4;
          const int c_sum = (int)(2u * (unsigned)g.kbase - (unsigned)ns + 1u - 2u * (unsigned)GH_GRP_MAGIC_BITS);
          ok = gh_sub_eq(U, V, hm, ns4, pix0, c_sum, pix);
        } else {
          const bool south = g.kind == GH_GRP_SOUTH;
          ok = gh_sub_polar(U, V, hm, south ? -2 : 2, g.kbase - 2, (int)((south ? (unsigned)npix : 0u) - (unsigned)GH_GRP_MAGIC_BITS), pix);
        }
        pix_out[(i * 8 + c) * 10 + s] = pix;
        ok_out[(i * 8 + c) * 10 + s] = ok ? 1 : 0;
      }
    }
  }
  return 0;
}

// S(o) = |C + p + o|^2 - |C|^2 in float against the double value, for the shell thresholds
extern "C" int gh_group_emulate_r2(double dx, const double *centres, long n, const float *sub_off_f, float *s_out)
{
  const float hg = (float)(dx * 1.7320508) + 1e-3f;
  for (long i = 0; i < n; ++i) {
    GhGroupExp g;
    gh_group_expand(centres[3 * i], centres[3 * i + 1], centres[3 * i + 2], hg, 256.f, 1.f, g);
    for (int c = 0; c < 8; ++c) {
      GhCellExp ce;
      const float px = (float)(((c & 1) - 0.5) * dx), py = (float)((((c >> 1) & 1) - 0.5) * dx), pz = (float)((((c >> 2) & 1) - 0.5) * dx);
      gh_group_recentre(g, px, py, pz, ce);
      for (int s = 0; s < 10; ++s) {
        const float ox = sub_off_f[s], oy = sub_off_f[10 + s], oz = sub_off_f[20 + s];
        s_out[(i * 8 + c) * 10 + s] = gh_cell_S(ce, ox, oy, oz, ox * ox + oy * oy + oz * oz);
      }
    }
  }
  return 0;
}
