/* The argument block of a run-time generated kernel (pysph_b200/codegen.py prepends this file
 * to the source it hands to NVRTC; b200sph.cu includes it).  Pool-wide arrays are indexed by
 * POOL index: d_idx / s_idx of the generated bodies are pool indices. */
#ifndef B200SPH_GENERIC_ARGS_H
#define B200SPH_GENERIC_ARGS_H
#define B200SPH_GENERIC_MAX_USER 16
typedef struct {
    double *f64[16];            /* x y z u v w rho h m x0 y0 z0 u0 v0 w0 rho0                  */
    float *f32[11];             /* p cs arho au av aw ax ay az dt_cfl dt_force                */
    unsigned int *u32[3];       /* gid tag pid                                                */
    double *user[B200SPH_GENERIC_MAX_USER]; /* b200sph_user_property arrays                   */
    const float4 *AB;           /* sorted packed records {A = (x y z rel. to the cell, h), B} */
    const unsigned char *stype; /* array id | ghost bit (0x08) of the particle at a sorted slot */
    const unsigned int *perm;   /* sorted slot -> pool index                                  */
    const unsigned int *cnt, *lst; /* the persistent neighbour lists, pair_list.cuh          */
    int capg;
    long long n;                /* sorted slots                                               */
    float cellx, celly, cellz, k2;
    int dest_type, real_only, phase;   /* phase: 0 initialize, 1 loop, 2 post_loop            */
    unsigned int src_mask;      /* bit a: array a is a source of some loop body of the launch */
    double t, dt;
    long long doff, dlo, dhi;   /* destinations [dlo, dhi) of the array that starts at doff   */
} b200sph_generic_args;
#endif
