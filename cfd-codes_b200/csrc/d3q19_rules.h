// d3q19_rules.h -- D3Q19 lattice and the wall/lid rule tables of the cavity, as X-macros.
//
// These tables restate WHAT the reference's HechtBC (Hecht & Harting non-equilibrium bounce-back)
// does per boundary class; reference = /root/reference/3D Lid-driven Cavity with Lattice
// Boltzmann Method/main.c ("3D:" below), lines 385-675.  Every rule is cell-local (the
// `fstar` parameter of HechtBC shadows the global and is always f2, 3D:385 vs 3D:70), so a rule is a
// short straight-line program over the 19 populations of ONE cell.  The kernels expand the tables
// into register-resident code; nothing here is executed by the CPU.
//
// Population numbering and velocities: 3D:36-38.  opp(q) = q+1 for odd q, q-1 for even q > 0.
//
// Boundary class of a cell: b? = 0 (strictly inside along that axis), 1 (coordinate == 0),
// 2 (coordinate == N-1);  class id = bx + 3*by + 9*bz  (0 = fluid interior, 26 wall classes).
#pragma once

#define LBM_CLS(bx, by, bz) ((bx) + 3 * (by) + 9 * (bz))

// ---- wall faces (not the lid) -------------------------------------------------------------------
// X(bx,by,bz,  un,on,                       normal unknown  f[un] = f[on]
//   a0,a1,a2, b0,b1,b2,                     N0 = 0.5*((f[a0]+f[a1]+f[a2]) - (f[b0]+f[b1]+f[b2]))
//   c0,c1,c2, d0,d1,d2,                     N1 = 0.5*((f[c0]+f[c1]+f[c2]) - (f[d0]+f[d1]+f[d2]))
//   u0,o0,s0, u1,o1,s1,                     f[u] = f[o] s N0     (s is the token + or -)
//   u2,o2,s2, u3,o3,s3)                     f[u] = f[o] s N1
// Both N terms are evaluated BEFORE any diagonal unknown is assigned; on the x = N-1 face N1 reads
// population 14, which is itself an unknown there (3D:421): the kernel loads the value the
// reference would find in its two-steps-old buffer (see lbm_kernels.cuh, "stale f14").
#define LBM_D3Q19_FACES(X)                                                                         \
    X(0, 0, 1, 5, 6, 1, 7, 13, 2, 14, 8, 3, 7, 14, 4, 13, 8, 9, 10, -, 16, 15, +, 11, 12, -, 18, 17, +)     /* z=0   3D:392-398 */ \
    X(0, 0, 2, 6, 5, 1, 7, 13, 2, 14, 8, 3, 7, 14, 4, 13, 8, 10, 9, +, 15, 16, -, 12, 11, +, 17, 18, -)     /* z=N-1 3D:401-407 */ \
    X(1, 0, 0, 1, 2, 3, 11, 17, 4, 18, 12, 5, 14, 11, 6, 17, 12, 13, 14, +, 7, 8, -, 9, 10, -, 15, 16, +)   /* x=0   3D:410-416 */ \
    X(2, 0, 0, 2, 1, 3, 11, 17, 4, 18, 12, 5, 14, 11, 6, 17, 12, 14, 13, -, 8, 7, +, 10, 9, +, 16, 15, -)   /* x=N-1 3D:419-425 */ \
    X(0, 1, 0, 3, 4, 1, 9, 15, 2, 16, 10, 5, 9, 16, 6, 15, 10, 7, 8, -, 14, 13, +, 11, 12, -, 17, 18, +)    /* y=0   3D:428-434 */

// The lid (y = N-1, class (0,2,0), 3D:437-445) has its own code in the kernel: wall density from
// the in-plane and outgoing populations, then the same pattern with rho*u_lid/3 and rho*(..)/6 terms.

// ---- edges ----------------------------------------------------------------------------------------
// X(bx,by,bz, coef, ta,tb,   temp = coef * (f[ta] - f[tb])
//   u0,s0, ... u8,s8)        f[u] = f[opp(u)] s temp, in this order; s in {P (+temp), M (-temp), Z (none)}
// Two of the nine unknowns of every edge are "buried" (their opposite is unknown too): they are
// only ever copied onto each other and keep the initial weight w[q] forever.
#define LBM_D3Q19_EDGES(X)                                                                         \
    X(0, 1, 1, 0.25, 1, 2, 3, Z, 7, M, 11, Z, 14, P, 17, Z, 5, Z, 9, M, 16, P, 18, Z)   /* 3D:452-463 */ \
    X(0, 1, 2, 0.25, 1, 2, 3, Z, 7, M, 11, Z, 14, P, 17, Z, 6, Z, 10, P, 12, Z, 15, M)  /* 3D:466-476 */ \
    X(0, 2, 1, 0.25, 1, 2, 4, Z, 8, P, 12, Z, 13, M, 18, Z, 5, Z, 9, M, 11, Z, 16, P)   /* 3D:479-489 */ \
    X(0, 2, 2, 0.25, 1, 2, 4, Z, 8, P, 12, Z, 13, M, 18, Z, 6, Z, 10, P, 15, M, 17, Z)  /* 3D:492-502 */ \
    X(1, 0, 1, 0.25, 3, 4, 1, Z, 7, M, 9, Z, 13, P, 15, Z, 5, Z, 11, M, 16, Z, 18, P)   /* 3D:505-515 */ \
    X(1, 0, 2, 0.25, 3, 4, 1, Z, 7, M, 9, Z, 13, P, 15, Z, 6, Z, 10, Z, 12, P, 17, M)   /* 3D:518-528 */ \
    X(2, 0, 1, 0.25, 3, 4, 2, Z, 8, M, 10, Z, 14, P, 16, Z, 5, Z, 9, Z, 11, M, 18, P)   /* 3D:531-541 */ \
    X(2, 0, 2, -0.25, 3, 4, 2, Z, 8, M, 10, Z, 14, P, 16, Z, 6, Z, 12, M, 15, Z, 17, P) /* 3D:544-554 */ \
    X(1, 1, 0, 0.25, 5, 6, 3, Z, 7, Z, 11, M, 14, Z, 17, P, 1, Z, 9, M, 13, Z, 15, P)   /* 3D:557-567 */ \
    X(1, 2, 0, 0.25, 5, 6, 4, Z, 8, Z, 12, P, 13, Z, 18, M, 1, Z, 7, Z, 9, M, 15, P)    /* 3D:570-580 */ \
    X(2, 1, 0, 0.25, 5, 6, 3, Z, 7, Z, 11, M, 14, Z, 17, P, 2, Z, 8, Z, 10, P, 16, M)   /* 3D:584-594 */ \
    X(2, 2, 0, 0.25, 5, 6, 4, Z, 8, Z, 12, P, 13, Z, 18, M, 2, Z, 10, P, 14, Z, 16, M)  /* 3D:598-608 */

// ---- corners ----------------------------------------------------------------------------------------
// X(bx,by,bz, u0..u5)   f[u] = f[opp(u)]  (the other six unknowns are buried constants)
#define LBM_D3Q19_CORNERS(X)                    \
    X(1, 2, 1, 1, 4, 5, 9, 13, 18)  /* 3D:613-618 */  \
    X(1, 1, 1, 1, 3, 5, 9, 7, 11)   /* 3D:621-626 */  \
    X(2, 2, 1, 2, 4, 5, 16, 8, 18)  /* 3D:629-634 */  \
    X(2, 1, 1, 2, 3, 5, 16, 14, 11) /* 3D:637-642 */  \
    X(1, 2, 2, 1, 4, 6, 15, 13, 12) /* 3D:645-650 */  \
    X(1, 1, 2, 1, 3, 6, 15, 7, 17)  /* 3D:653-658 */  \
    X(2, 2, 2, 2, 4, 6, 10, 8, 12)  /* 3D:661-666 */  \
    X(2, 1, 2, 2, 3, 6, 10, 14, 17) /* 3D:669-674 */
