// compat/mpi.h — in-process stand-in for the 10 MPI calls the reference uses.
//
// This image (and the GPU box) has no MPI. The reference's strategy sources
// ("Source Code/SparseMatrixFatVectorMultiply{RowWise,ColumnWise,NonZeroElement}.cpp")
// and main.cpp only use: MPI_Init, MPI_Finalize, MPI_Comm_size, MPI_Comm_rank,
// MPI_Barrier, MPI_Bcast, MPI_Gatherv, MPI_Reduce, MPI_Wtime, MPI_Abort with
// MPI_COMM_WORLD / MPI_INT / MPI_DOUBLE / MPI_SUM, always root 0, always blocking
// (SURVEY.md §2.3). Here a "rank" is a thread: compat_mpi::run(P, fn) starts P
// threads, each with a thread-local rank, and the collectives are memcpy/sum
// through shared slots fenced by a generation barrier. MPI_Reduce(SUM) adds the
// contributions in rank order 0,1,..,P-1 (deterministic; real MPI does not
// promise an order).
//
// If a real <mpi.h> is ever present, put it first on the include path instead.
#ifndef COMPAT_MPI_H
#define COMPAT_MPI_H

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_SUM 1
#define MPI_SUCCESS 0

namespace compat_mpi
{

struct World
{
    int size = 1;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long generation = 0;
    // slots published by ranks for the collective in flight
    std::vector<const void *> send;
    void *root_recv = nullptr;
    const int *root_counts = nullptr;
    const int *root_displs = nullptr;

    void barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        unsigned long gen = generation;
        if (++waiting == size)
        {
            waiting = 0;
            ++generation;
            cv.notify_all();
        }
        else
        {
            cv.wait(lk, [&] { return gen != generation; });
        }
    }
};

inline World g_world;
inline thread_local int t_rank = 0;

inline size_t type_size(MPI_Datatype t) { return t == MPI_DOUBLE ? sizeof(double) : sizeof(int); }

// Run fn(rank) on P rank-threads sharing one world; returns when all have finished.
inline void run(int P, const std::function<void(int)> &fn)
{
    if (P < 1)
        P = 1;
    g_world.size = P;
    g_world.waiting = 0;
    g_world.send.assign(P, nullptr);
    std::vector<std::thread> th;
    for (int r = 1; r < P; ++r)
        th.emplace_back([r, &fn] { t_rank = r; fn(r); });
    t_rank = 0;
    fn(0);
    for (auto &t : th)
        t.join();
    g_world.size = 1;
    t_rank = 0;
}

} // namespace compat_mpi

inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int *size)
{
    *size = compat_mpi::g_world.size;
    return MPI_SUCCESS;
}
inline int MPI_Comm_rank(MPI_Comm, int *rank)
{
    *rank = compat_mpi::t_rank;
    return MPI_SUCCESS;
}
inline int MPI_Barrier(MPI_Comm)
{
    compat_mpi::g_world.barrier();
    return MPI_SUCCESS;
}
inline double MPI_Wtime()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}
inline int MPI_Abort(MPI_Comm, int code)
{
    std::fflush(nullptr);
    std::_Exit(code);
}

inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm)
{
    compat_mpi::World &w = compat_mpi::g_world;
    if (w.size == 1)
        return MPI_SUCCESS;
    if (compat_mpi::t_rank == root)
        w.root_recv = buf;
    w.barrier();
    if (compat_mpi::t_rank != root && count > 0)
        std::memcpy(buf, w.root_recv, (size_t)count * compat_mpi::type_size(t));
    w.barrier();
    return MPI_SUCCESS;
}

inline int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype st,
                       void *recvbuf, const int *recvcounts, const int *displs,
                       MPI_Datatype, int root, MPI_Comm)
{
    compat_mpi::World &w = compat_mpi::g_world;
    const size_t es = compat_mpi::type_size(st);
    if (compat_mpi::t_rank == root)
    {
        w.root_recv = recvbuf;
        w.root_counts = recvcounts;
        w.root_displs = displs;
    }
    w.barrier();
    if (sendcount > 0)
        std::memcpy((char *)w.root_recv + (size_t)w.root_displs[compat_mpi::t_rank] * es,
                    sendbuf, (size_t)sendcount * es);
    w.barrier();
    return MPI_SUCCESS;
}

inline int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t,
                      MPI_Op, int root, MPI_Comm)
{
    compat_mpi::World &w = compat_mpi::g_world;
    w.send[compat_mpi::t_rank] = sendbuf;
    w.barrier();
    if (compat_mpi::t_rank == root)
    {
        if (t == MPI_DOUBLE)
        {
            double *out = (double *)recvbuf;
            const double *s0 = (const double *)w.send[0];
            for (int i = 0; i < count; ++i)
                out[i] = s0[i];
            for (int r = 1; r < w.size; ++r)
            {
                const double *s = (const double *)w.send[r];
                for (int i = 0; i < count; ++i)
                    out[i] += s[i];
            }
        }
        else
        {
            int *out = (int *)recvbuf;
            const int *s0 = (const int *)w.send[0];
            for (int i = 0; i < count; ++i)
                out[i] = s0[i];
            for (int r = 1; r < w.size; ++r)
            {
                const int *s = (const int *)w.send[r];
                for (int i = 0; i < count; ++i)
                    out[i] += s[i];
            }
        }
    }
    w.barrier();
    return MPI_SUCCESS;
}

#endif
