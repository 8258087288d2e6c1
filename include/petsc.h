/* petsc.h -- PETSc-SHAPED header of the p4b200 shim (NOT PETSc).
 *
 * It declares exactly the part of the PETSc API that the reference's structured-grid drivers use
 * (c/ch6/fish.c + c/ch6/poissonfunctions.c today), so that those files compile
 * UNCHANGED:   gcc -std=c99 -I include /root/reference/c/ch6/{fish.c,poissonfunctions.c} -lpetsc_p4b200
 * Behind it (p4pdes_b200/shim/petscshim.c) the solver objects are thin C structs whose numerical work
 * goes through the C ABI of include/p4b200.h into the sm_100a kernels:
 *
 *   DMDASNESSetFunctionLocal / DMDASNESSetJacobianLocal   the callback contract, kept bit-for-bit
 *        (c/ch6/fish.c:225-228; callbacks run on the HOST, once per solve / once per level)
 *   Mat (type "stencilcuda")   MatSetValuesStencil recognises the constant-coefficient 3/5/7-point
 *        stencil the callbacks insert (poissonfunctions.c:117-258); the operator is then applied
 *        matrix-free on the device
 *   PC  (type "mg")            p4b_mg_create_stencil / p4b_cg_solve: PCMG + Chebyshev/Jacobi + KSPCG
 *
 * Semantics follow PETSc's documented behaviour; deviations are listed in INTEGRATION.md.
 */
#ifndef P4B200_PETSC_SHIM_H_
#define P4B200_PETSC_SHIM_H_

#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PETSC_P4B200_SHIM 1

/* ---- basic types ---- */
typedef int PetscErrorCode;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef double PetscLogDouble;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
typedef enum { ENUM_DUMMY = 0 } PetscEnum;
typedef int MPI_Comm;

#define PETSC_COMM_WORLD ((MPI_Comm)1)
#define PETSC_COMM_SELF ((MPI_Comm)2)
#define PETSC_DECIDE (-1)
#define PETSC_DEFAULT (-2)
#define PETSC_PI 3.14159265358979323846264338327950288
#define PETSC_INFINITY (1.7976931348623157e308 / 4.0)

#define PetscExpReal(a) exp(a)
#define PetscSqrtReal(a) sqrt(a)
#define PetscAbsReal(a) fabs(a)
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))

/* ---- error handling: PetscCall returns on a non-zero code, SETERRQ reports and returns ---- */
PetscErrorCode PetscShimError(MPI_Comm comm, int line, const char *func, const char *file, PetscErrorCode code,
                              const char *msg);
#define PetscCall(...)                                   \
    do {                                                 \
        PetscErrorCode ierr_petsc_call_ = (__VA_ARGS__); \
        if (ierr_petsc_call_) return ierr_petsc_call_;   \
    } while (0)
#define SETERRQ(comm, code, msg) return PetscShimError(comm, __LINE__, __func__, __FILE__, code, msg)

/* ---- opaque objects ---- */
typedef struct _p_DM *DM;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_SNES *SNES;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_PetscRandom *PetscRandom;

typedef const char *SNESType;
typedef const char *KSPType;
typedef const char *PCType;
typedef const char *MatType;
#define SNESKSPONLY "ksponly"
#define SNESNEWTONLS "newtonls"
#define KSPCG "cg"
#define KSPGMRES "gmres"
#define KSPRICHARDSON "richardson"
#define KSPCHEBYSHEV "chebyshev"
#define PCMG "mg"
#define PCJACOBI "jacobi"
#define PCNONE "none"
#define MATSTENCILCUDA "stencilcuda"
#define MATSELLCUDA "sellcuda"

typedef enum { DM_BOUNDARY_NONE = 0, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_MIRROR, DM_BOUNDARY_PERIODIC } DMBoundaryType;
typedef enum { DMDA_STENCIL_STAR = 0, DMDA_STENCIL_BOX } DMDAStencilType;
typedef enum { NOT_SET_VALUES = 0, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;

/* DMDALocalInfo: field names and order as in PETSc (used at poissonfunctions.c:8-10,41-45,82-90) */
typedef struct {
    DM da;
    PetscInt dim, dof, sw;
    PetscInt mx, my, mz;       /* global number of grid points in each direction */
    PetscInt xs, ys, zs;       /* starting point of this processor, excluding ghosts */
    PetscInt xm, ym, zm;       /* number of grid points on this processor, excluding ghosts */
    PetscInt gxs, gys, gzs;    /* starting point of this processor including ghosts */
    PetscInt gxm, gym, gzm;    /* number of grid points on this processor including ghosts */
    DMBoundaryType bx, by, bz;
    DMDAStencilType st;
} DMDALocalInfo;

/* MatStencil: k, j, i, c in PETSc's order (poissonfunctions.c:121,126-137) */
typedef struct {
    PetscInt k, j, i, c;
} MatStencil;

/* the callback contract (fish.c:92-100,225-228; older spelling bratu2D.c:123-129) */
typedef PetscErrorCode DMDASNESFunctionFn(DMDALocalInfo *info, void *x, void *f, void *ctx);
typedef PetscErrorCode DMDASNESJacobianFn(DMDALocalInfo *info, void *x, Mat J, Mat Jpre, void *ctx);
typedef DMDASNESFunctionFn *DMDASNESFunction;
typedef DMDASNESJacobianFn *DMDASNESJacobian;

/* plugin registries (the PETSc idiom the north star names; used internally for -pc_type / -mat_type) */
PetscErrorCode MatRegister(const char name[], PetscErrorCode (*create)(Mat));
PetscErrorCode PCRegister(const char name[], PetscErrorCode (*create)(PC));

/* ---- system ---- */
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char file[], const char help[]);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char format[], ...);
PetscErrorCode PetscLogFlops(PetscLogDouble flops);

/* ---- options: PetscOptionsBegin/End bracket typed getters that share a prefix (fish.c:154-185) ---- */
PetscErrorCode PetscShimOptionsBegin(MPI_Comm comm, const char prefix[], const char title[], const char mansec[]);
PetscErrorCode PetscShimOptionsEnd(void);
#define PetscOptionsBegin(comm, prefix, title, mansec) \
    do {                                               \
        if (PetscShimOptionsBegin(comm, prefix, title, mansec)) return 98
#define PetscOptionsEnd()                      \
        if (PetscShimOptionsEnd()) return 98;  \
    } while (0)
PetscErrorCode PetscOptionsReal(const char opt[], const char text[], const char man[], PetscReal currentvalue,
                                PetscReal *value, PetscBool *set);
PetscErrorCode PetscOptionsInt(const char opt[], const char text[], const char man[], PetscInt currentvalue,
                               PetscInt *value, PetscBool *set);
PetscErrorCode PetscOptionsBool(const char opt[], const char text[], const char man[], PetscBool currentvalue,
                                PetscBool *value, PetscBool *set);
PetscErrorCode PetscOptionsEnum(const char opt[], const char text[], const char man[], const char *const *list,
                                PetscEnum currentvalue, PetscEnum *value, PetscBool *set);

/* ---- DM / DMDA ---- */
PetscErrorCode DMDACreate1d(MPI_Comm comm, DMBoundaryType bx, PetscInt M, PetscInt dof, PetscInt s,
                            const PetscInt lx[], DM *da);
PetscErrorCode DMDACreate2d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMDAStencilType st, PetscInt M,
                            PetscInt N, PetscInt m, PetscInt n, PetscInt dof, PetscInt s, const PetscInt lx[],
                            const PetscInt ly[], DM *da);
PetscErrorCode DMDACreate3d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz,
                            DMDAStencilType st, PetscInt M, PetscInt N, PetscInt P, PetscInt m, PetscInt n,
                            PetscInt p, PetscInt dof, PetscInt s, const PetscInt lx[], const PetscInt ly[],
                            const PetscInt lz[], DM *da);
PetscErrorCode DMSetApplicationContext(DM dm, void *ctx);
PetscErrorCode DMGetApplicationContext(DM dm, void *ctx);
PetscErrorCode DMSetFromOptions(DM dm);
PetscErrorCode DMSetUp(DM dm);
PetscErrorCode DMDASetUniformCoordinates(DM da, PetscReal xmin, PetscReal xmax, PetscReal ymin, PetscReal ymax,
                                         PetscReal zmin, PetscReal zmax);
PetscErrorCode DMGetBoundingBox(DM dm, PetscReal gmin[], PetscReal gmax[]);
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo *info);
PetscErrorCode DMGetGlobalVector(DM dm, Vec *g);
PetscErrorCode DMRestoreGlobalVector(DM dm, Vec *g);
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *g);
PetscErrorCode DMCreateMatrix(DM dm, Mat *mat);
PetscErrorCode DMDAVecGetArray(DM da, Vec vec, void *array);
PetscErrorCode DMDAVecRestoreArray(DM da, Vec vec, void *array);
PetscErrorCode DMDAVecGetArrayRead(DM da, Vec vec, void *array);
PetscErrorCode DMDAVecRestoreArrayRead(DM da, Vec vec, void *array);
PetscErrorCode DMDestroy(DM *dm);
PetscErrorCode DMDASNESSetFunctionLocal(DM dm, InsertMode imode, DMDASNESFunctionFn *func, void *ctx);
PetscErrorCode DMDASNESSetJacobianLocal(DM dm, DMDASNESJacobianFn *func, void *ctx);

/* ---- Vec ---- */
PetscErrorCode VecSet(Vec x, PetscScalar alpha);
PetscErrorCode VecSetRandom(Vec x, PetscRandom rctx);
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x);
PetscErrorCode VecScale(Vec x, PetscScalar alpha);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val);
PetscErrorCode VecGetSize(Vec x, PetscInt *size);
PetscErrorCode VecDuplicate(Vec v, Vec *newv);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r);
PetscErrorCode PetscRandomDestroy(PetscRandom *r);

/* ---- Mat ---- */
PetscErrorCode MatSetValuesStencil(Mat mat, PetscInt m, const MatStencil idxm[], PetscInt n, const MatStencil idxn[],
                                   const PetscScalar v[], InsertMode addv);
PetscErrorCode MatAssemblyBegin(Mat mat, MatAssemblyType type);
PetscErrorCode MatAssemblyEnd(Mat mat, MatAssemblyType type);
PetscErrorCode MatZeroEntries(Mat mat);
PetscErrorCode MatMult(Mat mat, Vec x, Vec y);
PetscErrorCode MatDestroy(Mat *mat);

/* ---- SNES / KSP ---- */
PetscErrorCode SNESCreate(MPI_Comm comm, SNES *snes);
PetscErrorCode SNESSetDM(SNES snes, DM dm);
PetscErrorCode SNESGetDM(SNES snes, DM *dm);
PetscErrorCode SNESSetType(SNES snes, SNESType type);
PetscErrorCode SNESGetKSP(SNES snes, KSP *ksp);
PetscErrorCode SNESSetFromOptions(SNES snes);
PetscErrorCode SNESSolve(SNES snes, Vec b, Vec x);
PetscErrorCode SNESGetSolution(SNES snes, Vec *x);
PetscErrorCode SNESGetIterationNumber(SNES snes, PetscInt *iter);
PetscErrorCode SNESDestroy(SNES *snes);
PetscErrorCode KSPSetType(KSP ksp, KSPType type);
PetscErrorCode KSPGetPC(KSP ksp, PC *pc);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode PCSetType(PC pc, PCType type);

/* ------------------------------------------------------------------------------------------------------
 * Declarations for c/ch7/minimal.c and c/ch5/pattern.c ("+ minimal.c", "+ pattern.c").
 * They let those files compile unchanged (oracle/_ref checks their callbacks against the oracle today);
 * the shim library does NOT implement the SNES-Newton / TS drivers behind them yet.
 * ------------------------------------------------------------------------------------------------------ */
#define PetscCoshReal(a) cosh(a)
#define PetscAcosReal(a) acos(a)
#define PetscSinReal(a) sin(a)
#define PetscCosReal(a) cos(a)
#define PetscPowReal(a, b) pow(a, b)

typedef struct _p_PetscObject *PetscObject;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_TS *TS;
typedef const char *TSType;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPIU_REAL ((MPI_Datatype)1)
#define MPIU_SUM ((MPI_Op)1)
#define MPIU_MIN ((MPI_Op)2)
#define MPIU_MAX ((MPI_Op)3)
#define TSARKIMEX "arkimex"
#define TSBEULER "beuler"
#define TSCN "cn"
#define TSBDF "bdf"
#define TSRK "rk"
typedef enum { TS_LINEAR = 0, TS_NONLINEAR } TSProblemType;
typedef enum { TS_EXACTFINALTIME_UNSPECIFIED = 0, TS_EXACTFINALTIME_STEPOVER, TS_EXACTFINALTIME_INTERPOLATE,
               TS_EXACTFINALTIME_MATCHSTEP } TSExactFinalTimeOption;
typedef struct { PetscScalar x, y; } DMDACoor2d;

PetscViewer PETSC_VIEWER_STDOUT_(MPI_Comm comm);
#define PETSC_VIEWER_STDOUT_WORLD PETSC_VIEWER_STDOUT_(PETSC_COMM_WORLD)
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
PetscErrorCode PetscObjectGetComm(PetscObject obj, MPI_Comm *comm);
PetscErrorCode PetscObjectGetTabLevel(PetscObject obj, PetscInt *tab);
PetscErrorCode PetscViewerASCIIAddTab(PetscViewer viewer, PetscInt tabs);
PetscErrorCode PetscViewerASCIISubtractTab(PetscViewer viewer, PetscInt tabs);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer viewer, const char format[], ...);
PetscErrorCode SNESMonitorSet(SNES snes, PetscErrorCode (*f)(SNES, PetscInt, PetscReal, void *), void *mctx,
                              PetscErrorCode (*monitordestroy)(void **));
PetscErrorCode DMGetLocalVector(DM dm, Vec *g);
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *g);
PetscErrorCode DMGlobalToLocalBegin(DM dm, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMGlobalToLocalEnd(DM dm, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMDASetFieldName(DM da, PetscInt nf, const char name[]);
PetscErrorCode DMDAGetCoordinateArray(DM da, void *xc);
PetscErrorCode DMDARestoreCoordinateArray(DM da, void *xc);

/* the TS callback contract (pattern.c:23-30,103-114) */
typedef PetscErrorCode (*DMDATSRHSFunctionLocal)(DMDALocalInfo *, PetscReal, void *, void *, void *);
typedef PetscErrorCode (*DMDATSRHSJacobianLocal)(DMDALocalInfo *, PetscReal, void *, Mat, Mat, void *);
typedef PetscErrorCode (*DMDATSIFunctionLocal)(DMDALocalInfo *, PetscReal, void *, void *, void *, void *);
typedef PetscErrorCode (*DMDATSIJacobianLocal)(DMDALocalInfo *, PetscReal, void *, void *, PetscReal, Mat, Mat, void *);
PetscErrorCode DMDATSSetRHSFunctionLocal(DM dm, InsertMode imode, DMDATSRHSFunctionLocal func, void *ctx);
PetscErrorCode DMDATSSetRHSJacobianLocal(DM dm, DMDATSRHSJacobianLocal func, void *ctx);
PetscErrorCode DMDATSSetIFunctionLocal(DM dm, InsertMode imode, DMDATSIFunctionLocal func, void *ctx);
PetscErrorCode DMDATSSetIJacobianLocal(DM dm, DMDATSIJacobianLocal func, void *ctx);
PetscErrorCode TSCreate(MPI_Comm comm, TS *ts);
PetscErrorCode TSSetProblemType(TS ts, TSProblemType type);
PetscErrorCode TSSetDM(TS ts, DM dm);
PetscErrorCode TSSetApplicationContext(TS ts, void *usrP);
PetscErrorCode TSSetType(TS ts, TSType type);
PetscErrorCode TSGetType(TS ts, TSType *type);
PetscErrorCode TSSetTime(TS ts, PetscReal t);
PetscErrorCode TSSetMaxTime(TS ts, PetscReal maxtime);
PetscErrorCode TSSetTimeStep(TS ts, PetscReal time_step);
PetscErrorCode TSSetExactFinalTime(TS ts, TSExactFinalTimeOption eftopt);
PetscErrorCode TSSetFromOptions(TS ts);
PetscErrorCode TSSolve(TS ts, Vec u);
PetscErrorCode TSDestroy(TS *ts);
/* + c/ch5/heat.c:72,83-84,117,133 */
PetscErrorCode TSMonitorSet(TS ts, PetscErrorCode (*monitor)(TS, PetscInt, PetscReal, Vec, void *), void *mctx,
                            PetscErrorCode (*mdestroy)(void **));
PetscErrorCode TSGetTime(TS ts, PetscReal *t);
PetscErrorCode TSGetMaxTime(TS ts, PetscReal *maxtime);
PetscErrorCode TSGetTimeStep(TS ts, PetscReal *dt);
PetscErrorCode TSGetDM(TS ts, DM *dm);

#ifdef __cplusplus
}
#endif
#endif /* P4B200_PETSC_SHIM_H_ */
