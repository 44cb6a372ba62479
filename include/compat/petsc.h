// compat/petsc.h — the PETSc surface the reference touches, as a tiny host stub.
//
// PETSc is a third-party comparison path in the reference (main.cpp:283-402,
// utils.cpp:8-29) and is OUT OF SCOPE for the B200 hot path.
// It is absent from this image, so this stub exists only so that the
// reference's utils.cpp / main.cpp compile and run unchanged. Each rank-thread
// owns its own Mat objects; only rank 0 inserts values (as main.cpp does), so
// rank 0's product is the full one. Not a PETSc re-implementation: dense
// row-major storage for MATDENSE, coordinate map for MATMPIAIJ, sequential
// product.
#ifndef COMPAT_PETSC_H
#define COMPAT_PETSC_H

#include <map>
#include <utility>
#include <vector>

#include <mpi.h>

typedef int PetscInt;
typedef double PetscScalar;
typedef int PetscErrorCode;
typedef int MatType;
typedef int InsertMode;
typedef int MatAssemblyType;
typedef int MatReuse;

#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_DECIDE (-1)
#define PETSC_DEFAULT (-2)
#define MATMPIAIJ 1
#define MATDENSE 2
#define INSERT_VALUES 1
#define MAT_FINAL_ASSEMBLY 0
#define MAT_INITIAL_MATRIX 0

struct _compat_Mat
{
    int rows = 0, cols = 0;
    MatType type = MATMPIAIJ;
    std::map<std::pair<int, int>, double> coo; // MATMPIAIJ
    std::vector<double> dense;                 // MATDENSE, row-major
};
typedef _compat_Mat *Mat;

inline double PetscRealPart(PetscScalar v) { return v; }
inline PetscErrorCode PetscInitialize(int *, char ***, const char *, const char *) { return 0; }
inline PetscErrorCode PetscFinalize() { return 0; }

inline PetscErrorCode MatCreate(MPI_Comm, Mat *m)
{
    *m = new _compat_Mat();
    return 0;
}
inline PetscErrorCode MatSetSizes(Mat m, PetscInt, PetscInt, PetscInt M, PetscInt N)
{
    m->rows = M;
    m->cols = N;
    return 0;
}
inline PetscErrorCode MatSetType(Mat m, MatType t)
{
    m->type = t;
    return 0;
}
inline PetscErrorCode MatSetUp(Mat m)
{
    if (m->type == MATDENSE)
        m->dense.assign((size_t)m->rows * m->cols, 0.0);
    return 0;
}
inline PetscErrorCode MatSetValue(Mat m, PetscInt i, PetscInt j, PetscScalar v, InsertMode)
{
    if (m->type == MATDENSE)
        m->dense[(size_t)i * m->cols + j] = v;
    else
        m->coo[{i, j}] = v;
    return 0;
}
inline PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
inline PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType) { return 0; }
inline PetscErrorCode MatProductCreate(Mat, Mat, Mat, Mat *) { return 0; }
inline PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse, double, Mat *C)
{
    Mat c = new _compat_Mat();
    c->rows = A->rows;
    c->cols = B->cols;
    c->type = MATDENSE;
    c->dense.assign((size_t)c->rows * c->cols, 0.0);
    if (!B->dense.empty())
        for (const auto &e : A->coo)
        {
            const int i = e.first.first, j = e.first.second;
            for (int k = 0; k < B->cols; ++k)
                c->dense[(size_t)i * c->cols + k] += e.second * B->dense[(size_t)j * B->cols + k];
        }
    *C = c;
    return 0;
}
inline PetscErrorCode MatCreateRedundantMatrix(Mat C, PetscInt, MPI_Comm, MatReuse, Mat *out)
{
    *out = new _compat_Mat(*C);
    return 0;
}
inline PetscErrorCode MatGetSize(Mat m, PetscInt *r, PetscInt *c)
{
    *r = m->rows;
    *c = m->cols;
    return 0;
}
inline PetscErrorCode MatGetValue(Mat m, PetscInt i, PetscInt j, PetscScalar *v)
{
    if (m->type == MATDENSE)
        *v = m->dense[(size_t)i * m->cols + j];
    else
    {
        auto it = m->coo.find({i, j});
        *v = it == m->coo.end() ? 0.0 : it->second;
    }
    return 0;
}
inline PetscErrorCode MatDestroy(Mat *m)
{
    delete *m;
    *m = nullptr;
    return 0;
}

#endif
