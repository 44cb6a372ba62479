// compat_main.cpp — process entry for reference-style drivers built on compat/mpi.h.
//
// TEST INFRASTRUCTURE / harness glue. The reference is started as
// "mpirun -np P ./main <k> <file.mtx>" (scripts/mpi.sub:97). Without MPI in the
// image, the driver's own main() is renamed to compat_user_main at compile time
// (-Dmain=compat_user_main) and this file provides main(): P comes from
// "-np P" in front of the program arguments or from COMPAT_MPI_NP (default 1),
// and every rank-thread runs compat_user_main with the same argv.
#include <cstdlib>
#include <cstring>
#include <vector>

#include <mpi.h>

int compat_user_main(int argc, char *argv[]);

int main(int argc, char *argv[])
{
    int P = 1;
    if (const char *e = std::getenv("COMPAT_MPI_NP"))
        P = std::atoi(e);
    std::vector<char *> args;
    args.push_back(argv[0]);
    for (int i = 1; i < argc; ++i)
    {
        if (std::strcmp(argv[i], "-np") == 0 && i + 1 < argc)
        {
            P = std::atoi(argv[++i]);
            continue;
        }
        args.push_back(argv[i]);
    }
    int rc = 0;
    compat_mpi::run(P, [&](int rank) {
        std::vector<char *> mine(args); // each rank gets its own argv array
        mine.push_back(nullptr);
        int r = compat_user_main((int)args.size(), mine.data());
        if (rank == 0)
            rc = r;
    });
    return rc;
}
