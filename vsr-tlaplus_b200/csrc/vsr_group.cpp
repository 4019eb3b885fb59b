/*
 * vsr_group.cpp — the rank group of a multi-GPU job (vsr_group.h): shared-memory barrier and small all-gather.
 * C ABI: vsr_group_open / vsr_group_open_local / vsr_group_close / vsr_group_barrier / vsr_group_allgather /
 * vsr_group_abort (include/vsr_b200.h).
 */
#include "vsr_group.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <chrono>
#include <new>

#include "../../include/vsr_b200.h"

using namespace vsr;

static const uint32_t GROUP_MAGIC = 0x56535247u; /* "VSRG" */

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void relax(unsigned spins) {
    if (spins < 2000) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    } else if (spins < 20000) {
        sched_yield();
    } else {
        struct timespec ts = {0, 50000}; /* 50 us: ranks that share cores with the waiters (tests) must get to run */
        nanosleep(&ts, nullptr);
    }
}

static int group_fail(VsrGroup* g, const char* what) {
    snprintf(g->last_error, sizeof g->last_error, "rank %d/%d: %s", g->rank, g->world, what);
    return VSR_RC_SYSTEM;
}

extern "C" {

int vsr_group_open(const char* name, int rank, int world, double timeout_s, VsrGroup** out, char* err, size_t errcap) {
    auto fail = [&](const char* msg) {
        if (err && errcap) snprintf(err, errcap, "vsr_group_open(%s, rank %d of %d): %s%s%s", name ? name : "?", rank, world, msg, errno ? ": " : "",
                                    errno ? strerror(errno) : "");
        return VSR_RC_SYSTEM;
    };
    errno = 0;
    if (!name || name[0] != '/' || !out) return fail("the name must start with '/'");
    if (world < 1 || world > VSR_GROUP_MAX_WORLD || rank < 0 || rank >= world) return fail("world must be 1..8 and 0 <= rank < world");
    if (timeout_s <= 0) timeout_s = 120.0;
    const size_t bytes = (sizeof(GroupShm) + 4095) & ~(size_t)4095;
    int fd = -1;
    const double t0 = now_s();
    if (rank == 0) {
        shm_unlink(name); /* a stale block of a crashed job with the same name */
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) return fail("shm_open(create)");
        if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); return fail("ftruncate"); }
    } else {
        for (unsigned spins = 0;; spins++) { /* wait for rank 0 to create and size it */
            fd = shm_open(name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes) break;
                close(fd);
                fd = -1;
            }
            if (now_s() - t0 > timeout_s) { errno = 0; return fail("timed out waiting for rank 0 to create the group"); }
            relax(20000 + spins);
        }
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail("mmap");
    GroupShm* shm = (GroupShm*)p;
    if (rank == 0) {
        /* fresh pages are zero; publish */
        shm->world.store((uint32_t)world);
        shm->magic.store(GROUP_MAGIC, std::memory_order_release);
    } else {
        for (unsigned spins = 0; shm->magic.load(std::memory_order_acquire) != GROUP_MAGIC; spins++) {
            if (now_s() - t0 > timeout_s) { munmap(p, bytes); errno = 0; return fail("timed out waiting for rank 0 to initialise the group"); }
            relax(spins);
        }
        if (shm->world.load() != (uint32_t)world) { munmap(p, bytes); errno = 0; return fail("world size differs from rank 0's (a stale group of another job?)"); }
    }
    shm->attached.fetch_add(1);
    for (unsigned spins = 0; shm->attached.load() < (uint32_t)world; spins++) {
        if (now_s() - t0 > timeout_s) { munmap(p, bytes); if (rank == 0) shm_unlink(name); errno = 0; return fail("timed out waiting for the other ranks to attach"); }
        relax(spins);
    }
    VsrGroup* g = new VsrGroup();
    g->shm = shm;
    g->rank = rank;
    g->world = world;
    g->map_bytes = bytes;
    snprintf(g->name, sizeof g->name, "%s", name);
    if (rank == 0) shm_unlink(name); /* everybody has it mapped: the name can go, the memory lives until the last unmap */
    *out = g;
    return 0;
}

int vsr_group_open_local(int world, VsrGroup** out_handles) {
    if (world < 1 || world > VSR_GROUP_MAX_WORLD || !out_handles) return VSR_RC_ERROR;
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(GroupShm)) != 0) return VSR_RC_SYSTEM;
    memset(mem, 0, sizeof(GroupShm));
    GroupShm* shm = new (mem) GroupShm;
    shm->world.store((uint32_t)world);
    shm->attached.store((uint32_t)world);
    shm->magic.store(GROUP_MAGIC);
    std::atomic<int>* refs = new std::atomic<int>(world);
    for (int r = 0; r < world; r++) {
        VsrGroup* g = new VsrGroup();
        g->shm = shm;
        g->rank = r;
        g->world = world;
        g->local = true;
        g->local_refs = refs;
        out_handles[r] = g;
    }
    return 0;
}

void vsr_group_close(VsrGroup* g) {
    if (!g) return;
    if (g->local) {
        if (g->local_refs->fetch_sub(1) == 1) {
            free(g->shm);
            delete g->local_refs;
        }
    } else if (g->shm) {
        munmap(g->shm, g->map_bytes);
    }
    delete g;
}

void vsr_group_abort(VsrGroup* g) {
    if (g && g->shm) g->shm->abort.store(1);
}

int vsr_group_rank(const VsrGroup* g) { return g->rank; }
int vsr_group_world(const VsrGroup* g) { return g->world; }
const char* vsr_group_last_error(const VsrGroup* g) { return g->last_error; }
void vsr_group_set_timeout(VsrGroup* g, double seconds) { if (seconds > 0) g->timeout_s = seconds; }

int vsr_group_barrier(VsrGroup* g) {
    GroupShm* s = g->shm;
    if (s->abort.load()) return group_fail(g, "another rank aborted the job");
    if (g->world == 1) return 0;
    const uint32_t gen = s->generation.load(std::memory_order_acquire);
    if (s->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)g->world) {
        s->arrived.store(0, std::memory_order_relaxed);
        s->generation.store(gen + 1, std::memory_order_release);
        return 0;
    }
    const double t0 = now_s();
    for (unsigned spins = 0; s->generation.load(std::memory_order_acquire) == gen; spins++) {
        if (s->abort.load(std::memory_order_relaxed)) return group_fail(g, "another rank aborted the job");
        if ((spins & 1023) == 1023 && now_s() - t0 > g->timeout_s) {
            s->abort.store(1);
            return group_fail(g, "barrier timed out (a rank died or hangs)");
        }
        relax(spins);
    }
    return 0;
}

int vsr_group_allgather(VsrGroup* g, const void* mine, size_t bytes, void* all_out) {
    if (bytes > VSR_GROUP_MSG_BYTES) return group_fail(g, "all-gather message too large");
    GroupShm* s = g->shm;
    const int buf = (int)(g->seq++ & 1);
    memcpy(s->slots[buf][g->rank], mine, bytes);
    const int rc = vsr_group_barrier(g);
    if (rc) return rc;
    for (int r = 0; r < g->world; r++) memcpy((uint8_t*)all_out + (size_t)r * bytes, s->slots[buf][r], bytes);
    /* no second barrier: the next all-gather writes the other buffer, and nobody can write this one again before
       everybody has passed the next barrier, i.e. has finished reading here */
    return 0;
}

} /* extern "C" */
