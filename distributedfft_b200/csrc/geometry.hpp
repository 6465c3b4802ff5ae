// geometry.hpp — host-only partition / layout arithmetic of a plan (no CUDA).
//
// Restates the split rule of the reference's initFFT:
//   slab   /root/reference/src/slab/default/mpicufft_slab.cpp:112-128   (x split -> y split)
//   z_yx   /root/reference/src/slab/z_then_yx/mpicufft_slab_z_then_yx.cpp (x split -> z split)
//   pencil /root/reference/src/pencil/mpicufft_pencil.cpp:84-110         (pidx = i*P2 + j)
// size[p] = n/parts + (p < n%parts); start = prefix sum.
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

#include "../../include/dfft.h"

namespace dfft {

struct Split {
    std::vector<size_t> size, start;
    void make(size_t n, size_t parts) {
        size.assign(parts, n / parts);
        for (size_t p = 0; p < n % parts; ++p) size[p]++;
        start.assign(parts, 0);
        for (size_t p = 1; p < parts; ++p) start[p] = start[p - 1] + size[p - 1];
    }
};

struct Geometry {
    int decomp = 0, transform = 0;
    size_t nx = 0, ny = 0, nz = 0, nzc = 0;  // nzc: z extent of every complex array
    int P = 1, P1 = 1, P2 = 1;
    // x split of the input (over P1, or over P for both slab sequences)
    Split sx;
    // pencil: y split of the input over P2;  z split of the spectrum over P2 (z_yx: over P)
    Split sy, sz;
    // y split of the output over P1 (slab zy_x: over P)
    Split oy;

    // returns false on an invalid partition
    bool init(int decomp_, int transform_, size_t nx_, size_t ny_, size_t nz_, size_t p1, size_t p2, int nranks) {
        decomp = decomp_; transform = transform_;
        nx = nx_; ny = ny_; nz = nz_;
        nzc = (transform == DFFT_C2C) ? nz : nz / 2 + 1;
        P = nranks;
        if (decomp == DFFT_SLAB_ZY_THEN_X) { P1 = P; P2 = 1; }
        else if (decomp == DFFT_SLAB_Z_THEN_YX) { P1 = P; P2 = 1; }  // P1 only used for the x split
        else if (decomp == DFFT_PENCIL) { P1 = int(p1); P2 = int(p2); }
        else return false;
        if (P1 < 1 || P2 < 1 || size_t(P1) * size_t(P2) != size_t(P)) return false;
        if (nx == 0 || ny == 0 || nz == 0) return false;
        sx.make(nx, P1);
        if (decomp == DFFT_SLAB_Z_THEN_YX) {
            sy.make(ny, 1);
            sz.make(nzc, P);
            oy.make(ny, 1);
        } else {
            sy.make(ny, P2);
            sz.make(nzc, P2);
            oy.make(ny, P1);
        }
        return true;
    }
    int pi(int rank) const { return decomp == DFFT_PENCIL ? rank / P2 : rank; }
    int pj(int rank) const { return decomp == DFFT_PENCIL ? rank % P2 : 0; }

    // which: 0 input, 1 after z, 2 after z and y, 3 output
    void layout(int rank, int which, size_t size[3], size_t start[3]) const {
        const int i = pi(rank), j = pj(rank);
        if (decomp == DFFT_SLAB_Z_THEN_YX) {
            if (which <= 1) {
                size[0] = sx.size[i]; size[1] = ny; size[2] = which == 0 ? nz : nzc;
                start[0] = sx.start[i]; start[1] = 0; start[2] = 0;
            } else {
                size[0] = nx; size[1] = ny; size[2] = sz.size[rank];
                start[0] = 0; start[1] = 0; start[2] = sz.start[rank];
            }
            return;
        }
        switch (which) {
            case 0:
            case 1:
                size[0] = sx.size[i]; size[1] = sy.size[j]; size[2] = which == 0 ? nz : nzc;
                start[0] = sx.start[i]; start[1] = sy.start[j]; start[2] = 0;
                break;
            case 2:
                size[0] = sx.size[i]; size[1] = ny; size[2] = sz.size[j];
                start[0] = sx.start[i]; start[1] = 0; start[2] = sz.start[j];
                break;
            default:
                size[0] = nx; size[1] = oy.size[i]; size[2] = sz.size[j];
                start[0] = 0; start[1] = oy.start[i]; start[2] = sz.start[j];
                break;
        }
    }
    // complex elements of the largest per-rank stage array (the reference's domainsize / sizeof(C_t),
    // mpicufft_slab.cpp:132)
    size_t domain_elems(int rank) const {
        size_t m = 0;
        for (int w = 1; w <= 3; ++w) {
            size_t s[3], o[3];
            layout(rank, w, s, o);
            m = std::max(m, s[0] * s[1] * s[2]);
        }
        return m;
    }
};

}  // namespace dfft
