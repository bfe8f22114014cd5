/*
 * oracle/mpi_shim/mpi.h -- TEST SCAFFOLDING: a tiny fork + shared-memory stand-in for the handful
 * of MPI calls the reference's 2D program makes (no MPI is installed in this image).  It exists only
 * so that oracle/build_ref.py can compile the UNMODIFIED reference
 *   /root/reference/2D Lid-driven Cavity with Lattice Boltzmann Method/main.cpp
 * into oracle/_ref/lbm2d_ref_* and the in-place oracle (oracle/lbm2d_oracle.c, scheme B) can be
 * pinned against it.  Not part of the product.
 *
 * Semantics: MPI_Init mmaps a MAP_SHARED region and fork()s SHIM_NP-1 children (environment
 * variable SHIM_NP, default 3).  MPI_Isend copies eagerly into a FIFO ring owned by (destination,
 * tag); MPI_Irecv only records the request; MPI_Waitall completes receives in FIFO order, which
 * gives MPI's same-(source,tag) message ordering -- every (destination, tag) pair of the program has
 * exactly one source.  MPI_Allreduce(SUM) adds the per-rank slots in rank order.
 */
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
struct MPI_Status { int dummy; };
struct MPI_Request { int kind; void *buf; int count; int tag; };  /* kind 0 = null, 2 = pending recv */
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_SUM 1
#define MPI_REQUEST_NULL (MPI_Request{0, nullptr, 0, 0})

#define SHIM_MAXP 128
#define SHIM_TAGS 3
#define SHIM_DEPTH 16
#define SHIM_MAXCOUNT 16384

struct shim_ring {
    volatile long wr, rd;
    int count[SHIM_DEPTH];
    double data[SHIM_DEPTH][SHIM_MAXCOUNT];
};
struct shim_shared {
    volatile int bar_count, bar_sense;
    double slots[SHIM_MAXP];
    shim_ring ring[SHIM_MAXP][SHIM_TAGS];
};

static shim_shared *shim_g = nullptr;
static int shim_rank = 0, shim_np = 1, shim_local_sense = 0;
static pid_t shim_children[SHIM_MAXP];

static inline int MPI_Barrier(MPI_Comm) {
    shim_local_sense = !shim_local_sense;
    if (__sync_add_and_fetch(&shim_g->bar_count, 1) == shim_np) {
        shim_g->bar_count = 0;
        __sync_synchronize();
        shim_g->bar_sense = shim_local_sense;
    } else {
        while (shim_g->bar_sense != shim_local_sense) sched_yield();
    }
    __sync_synchronize();
    return 0;
}

static inline int MPI_Init(int *, char ***) {
    const char *e = getenv("SHIM_NP");
    shim_np = e ? atoi(e) : 3;
    if (shim_np < 1 || shim_np > SHIM_MAXP) { fprintf(stderr, "SHIM_NP out of range\n"); exit(2); }
    shim_g = (shim_shared *)mmap(nullptr, sizeof(shim_shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shim_g == MAP_FAILED) { perror("mmap"); exit(2); }
    fflush(stdout);
    for (int r = 1; r < shim_np; r++) {
        pid_t pid = fork();
        if (pid == 0) { shim_rank = r; break; }
        shim_children[r] = pid;
    }
    return 0;
}
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = shim_rank; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = shim_np; return 0; }
static inline double MPI_Wtime() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static inline int MPI_Isend(const void *buf, int count, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request *req) {
    if (count > SHIM_MAXCOUNT) { fprintf(stderr, "shim: message too large\n"); exit(2); }
    shim_ring *q = &shim_g->ring[dest][tag];
    while (q->wr - q->rd >= SHIM_DEPTH) sched_yield();
    const int slot = (int)(q->wr % SHIM_DEPTH);
    memcpy(q->data[slot], buf, (size_t)count * sizeof(double));
    q->count[slot] = count;
    __sync_synchronize();
    q->wr = q->wr + 1;
    *req = MPI_REQUEST_NULL;
    return 0;
}
static inline int MPI_Irecv(void *buf, int count, MPI_Datatype, int, int tag, MPI_Comm, MPI_Request *req) {
    *req = MPI_Request{2, buf, count, tag};
    return 0;
}
static inline int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *) {
    for (int i = 0; i < n; i++) {
        if (reqs[i].kind != 2) continue;
        shim_ring *q = &shim_g->ring[shim_rank][reqs[i].tag];
        while (q->rd >= q->wr) sched_yield();
        __sync_synchronize();
        const int slot = (int)(q->rd % SHIM_DEPTH);
        memcpy(reqs[i].buf, q->data[slot], (size_t)reqs[i].count * sizeof(double));
        __sync_synchronize();
        q->rd = q->rd + 1;
        reqs[i] = MPI_REQUEST_NULL;
    }
    return 0;
}
static inline int MPI_Allreduce(const void *in, void *out, int count, MPI_Datatype, MPI_Op, MPI_Comm) {
    if (count != 1) { fprintf(stderr, "shim: Allreduce count != 1\n"); exit(2); }
    shim_g->slots[shim_rank] = *(const double *)in;
    MPI_Barrier(0);
    double s = 0.0;
    for (int r = 0; r < shim_np; r++) s += shim_g->slots[r];
    *(double *)out = s;
    MPI_Barrier(0);
    return 0;
}
static inline int MPI_Finalize() {
    MPI_Barrier(0);
    fflush(stdout);
    if (shim_rank == 0) {
        for (int r = 1; r < shim_np; r++) { int st; waitpid(shim_children[r], &st, 0); }
    } else {
        _exit(0);
    }
    return 0;
}
#endif
