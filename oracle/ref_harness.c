/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Drives the UNMODIFIED reference solver
 *   /root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c
 * from the outside.  build_ref.py pipes that file through a parameter patch of
 * its `#define` lines (3D:7-15; they are macros, so -D cannot override them)
 * into a temporary file and compiles THIS file with -DREF_SRC="<tmp>"; the
 * reference text itself is never written into the repository.  Outputs go to
 * oracle/_ref/ only (git-ignored, travels to the GPU box as a built binary).
 *
 * Sub-commands
 *   steps <n> <out.bin> [threads]   run n iterations of the reference time loop
 *                                   (3D:94-100 / 3D:106-123: macrocollideandstream();
 *                                   HechtBC(f2); swap) and dump the newest lattice
 *                                   (NQ doubles, LU4 order 3D:48).  Prints the
 *                                   diff() (3D:678) of the last iteration.
 *   time <n> [threads]              the same loop, timed with omp_get_wtime;
 *                                   prints one JSON line (used as cpu_baseline
 *                                   kind "reference" by bench.py).
 *   bench <iters> <warm> <steps> [threads]
 *                                   `warm` untimed + `steps` timed bench steps of `iters`
 *                                   iterations each (bench.py --impl reference); JSON line.
 *   sample <seconds> <warm> [threads]
 *                                   `warm` untimed iterations, then whole iterations until at least
 *                                   <seconds> of wall time have passed (>= 3): the bounded CPU
 *                                   sample of bench.py's cpu_baseline; JSON line.
 *   main                            runs the reference's own main() (3D:74-132)
 *                                   in the current directory (banner, residual
 *                                   blocks, Solution_*.dat).
 */
#define main ref_main
#include REF_SRC
#undef main

#include <string.h>

static void set_threads(int nt) {
    /* allocate() -> initOMP() (3D:209-212) pins THREADS; every later
     * `#pragma omp parallel for` uses the current ICV, so override it after. */
    if (nt > 0) omp_set_num_threads(nt);
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s steps <n> <out.bin> [threads] | time <n> [threads] | bench <iters> <warm> <steps> [threads] | main\n", argv[0]);
        return 2;
    }
    if (strcmp(argv[1], "main") == 0) return ref_main();

    if (strcmp(argv[1], "steps") == 0 && argc >= 4) {
        int n = atoi(argv[2]);
        int nt = argc > 4 ? atoi(argv[4]) : 0;
        allocate();
        initialize();
        set_threads(nt);
        double d = 0.0;
        for (int t = 0; t < n; t++) {
            macrocollideandstream();
            HechtBC(f2);
            if (t == n - 1) d = diff(f2, f1, NQ);
            swap(&f1, &f2);
        }
        FILE *o = fopen(argv[3], "wb");
        if (!o) { perror("fopen"); return 1; }
        size_t nq = (size_t)NQ;
        if (fwrite(f1, sizeof(double), nq, o) != nq) { perror("fwrite"); return 1; }
        fclose(o);
        printf("{\"N\": %d, \"steps\": %d, \"last_diff\": %.17g}\n", N, n, d);
        return 0;
    }

    if (strcmp(argv[1], "time") == 0 && argc >= 3) {
        int n = atoi(argv[2]);
        int nt = argc > 3 ? atoi(argv[3]) : 0;
        allocate();
        initialize();
        set_threads(nt);
        /* one untimed iteration (page faults of f2, iteration-0 special case) */
        macrocollideandstream();
        HechtBC(f2);
        swap(&f1, &f2);
        double t0 = omp_get_wtime();
        for (int t = 0; t < n; t++) {
            macrocollideandstream();
            HechtBC(f2);
            swap(&f1, &f2);
        }
        double dt = omp_get_wtime() - t0;
        double cells = (double)N * (double)N * (double)N;
        printf("{\"N\": %d, \"steps\": %d, \"seconds\": %.6f, \"mlups\": %.6f, \"threads\": %d}\n",
               N, n, dt, cells * n / dt / 1e6, omp_get_max_threads());
        return 0;
    }
    if (strcmp(argv[1], "bench") == 0 && argc >= 5) {
        int iters = atoi(argv[2]), warm = atoi(argv[3]), steps = atoi(argv[4]);
        int nt = argc > 5 ? atoi(argv[5]) : 0;
        allocate();
        initialize();
        set_threads(nt);
        for (int t = 0; t < warm * iters; t++) {
            macrocollideandstream();
            HechtBC(f2);
            swap(&f1, &f2);
        }
        double t0 = omp_get_wtime();
        for (int t = 0; t < steps * iters; t++) {
            macrocollideandstream();
            HechtBC(f2);
            swap(&f1, &f2);
        }
        double dt = omp_get_wtime() - t0;
        double cells = (double)N * (double)N * (double)N;
        printf("{\"N\": %d, \"iters_per_step\": %d, \"steps\": %d, \"warmup\": %d, \"seconds\": %.6f, "
               "\"mlups\": %.6f, \"threads\": %d}\n",
               N, iters, steps, warm, dt, cells * steps * iters / dt / 1e6, omp_get_max_threads());
        return 0;
    }
    if (strcmp(argv[1], "sample") == 0 && argc >= 4) {
        double want = atof(argv[2]);
        int warm = atoi(argv[3]);
        int nt = argc > 4 ? atoi(argv[4]) : 0;
        allocate();
        initialize();
        set_threads(nt);
        for (int t = 0; t < warm; t++) {
            macrocollideandstream();
            HechtBC(f2);
            swap(&f1, &f2);
        }
        int n = 0;
        double t0 = omp_get_wtime(), dt = 0.0;
        while (n < 3 || dt < want) {
            macrocollideandstream();
            HechtBC(f2);
            swap(&f1, &f2);
            n++;
            dt = omp_get_wtime() - t0;
        }
        double cells = (double)N * (double)N * (double)N;
        printf("{\"N\": %d, \"iters_per_step\": 1, \"steps\": %d, \"warmup\": %d, \"seconds\": %.6f, "
               "\"mlups\": %.6f, \"threads\": %d}\n",
               N, n, warm, dt, cells * n / dt / 1e6, omp_get_max_threads());
        return 0;
    }
    fprintf(stderr, "bad arguments\n");
    return 2;
}
