// main.cpp -- `yaha_b200`: drop-in for the alignment mode of the yaha 0.1.83 command line
// (Main.c:187-665).  Prints the same banner and resource-usage line.
#include <malloc.h>
#include <stdlib.h>
#include <sys/resource.h>
#include <time.h>
#include "host.hpp"

int main(int argc, char **argv)
{
    fprintf(stderr, "YAHA version 0.1.83 (yaha_b200: B200-native alignment hot path)\n");
    // batches allocate and release tens of MB at a time: keep the heap instead of trimming it back to
    // the kernel (and re-faulting the pages) after every batch
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
    yh::Args A;
    if (yh::parseArgs(argc, argv, A) != 0) return 0;
    time_t t0 = time(nullptr);
    int rc = A.query ? yh::runQueries(A) : yh::runIndex(A);
    time_t t1 = time(nullptr);
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    double u = ru.ru_utime.tv_sec + ru.ru_utime.tv_usec / 1e6, s = ru.ru_stime.tv_sec + ru.ru_stime.tv_usec / 1e6;
    fprintf(stderr, "Operation on %s used %.3fS User, %.3fS System, Total: %.3fS in %ldS wall time.\n",
            A.query ? A.qfile.c_str() : A.gfile.c_str(), u, s, u + s, (long)(t1 - t0));
    return rc;
}
