/*
 * vsr_group.h — internal: the ranks of one multi-GPU job on one node (one per GPU; processes under torchrun, or threads of
 * one process under `vsrmc -gpus N`).  All the host-side coordination the sharded BFS needs — a barrier and an all-gather
 * of one small message per rank, a few times per wavefront — goes through one block of shared memory (POSIX shm between
 * processes, the heap between threads): a round trip costs about a microsecond, against tens of microseconds for a
 * collective launched from Python.  The states themselves never pass through here: the expand kernel stores them into
 * the owner's inbox over NVLink (vsr_gpu.cuh push_records).  No CUDA in this file (it is tested on CPU).
 */
#ifndef VSR_GROUP_H
#define VSR_GROUP_H

#include <stddef.h>
#include <stdint.h>

#include <atomic>

#define VSR_GROUP_MAX_WORLD 8
#define VSR_GROUP_MSG_BYTES 256

namespace vsr {

struct GroupShm {
    std::atomic<uint32_t> magic;      /* set by the creator once the block is initialised */
    std::atomic<uint32_t> world;
    std::atomic<uint32_t> attached;   /* ranks that have mapped the block */
    std::atomic<uint32_t> arrived;    /* barrier: ranks that have arrived in the current generation */
    std::atomic<uint32_t> generation;
    std::atomic<int32_t> abort;       /* a rank failed: every wait returns an error instead of hanging */
    uint32_t _pad[10];
    alignas(64) uint8_t slots[2][VSR_GROUP_MAX_WORLD][VSR_GROUP_MSG_BYTES]; /* all-gather payload, double-buffered */
};

} // namespace vsr

struct VsrGroup {
    vsr::GroupShm* shm = nullptr;
    int rank = 0, world = 1;
    bool local = false;        /* threads of one process: shm is heap memory shared by the world handles */
    std::atomic<int>* local_refs = nullptr;
    uint32_t seq = 0;          /* all-gathers done (selects the slot buffer) */
    double timeout_s = 600.0;
    size_t map_bytes = 0;
    char name[96] = {0};
    char last_error[160] = {0};
};

#endif
