/*
 * q1_internal.h -- what the translation units of libq1phys share besides the public ABI
 * (include/q1phys.h): the policy handle, and the view of an env handle the fused policy + env kernels
 * (q1_actor.cu) need.  Nothing here is exported.
 */
#pragma once

#include "../../include/q1phys.h"
#include "q1_tick.cuh"

#include <string>

int q1_set_error(int code, const std::string &msg); /* q1phys.cu: thread-local last error */

struct q1_policy {
    int device = 0;
    int num_keys = 4;
    int sm_count = 148;
    unsigned char *image = nullptr; /* device copy of the shared-memory weight image (q1_actor.cu) */
};

/* The part of a q1_env a kernel outside q1phys.cu needs. */
struct q1_env_view {
    q1::Params P;
    int device;
    bool stamps, track;
    uint64_t ticks;        /* ticks executed so far (position of the policy noise stream) */
    int sm_count;
};
/* Fills `out`; `on_caller_stream` marks the handle as driven on a caller's stream (see
 * q1_env::caller_streams_used in q1phys.cu).  Returns Q1_OK or Q1_EINVAL. */
int q1_env_get_view(q1_env *env, bool on_caller_stream, q1_env_view *out);
/* The handle's private stream and a device scratch buffer of at least `bytes` (the *_host entry
 * points' staging area); waits for caller-stream work first, like every *_host call. */
int q1_env_host_scratch(q1_env *env, size_t bytes, void **scratch, void **stream);
