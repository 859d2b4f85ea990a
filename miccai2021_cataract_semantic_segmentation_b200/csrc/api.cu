// Error plumbing, version and device queries of the b200seg C ABI.
#include "b200seg.h"
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static thread_local char g_err[512] = "";

void b200seg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int b200seg_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
B200segTuning& b200seg_tuning() {
    static B200segTuning t = {env_int("B200SEG_INTERLEAVE", 1), env_int("B200SEG_STATS_VARIANT", 0),
                              env_int("B200SEG_EMIT_PATH", 0), env_int("B200SEG_SORT_MATCH", 2), env_int("B200SEG_SORT_PATH", 0),
                              env_int("B200SEG_DBG", 0), env_int("B200SEG_PDL", 1)};
    return t;
}
extern "C" int b200seg_set_tuning(const char* key, int32_t value) {
    B200segTuning& t = b200seg_tuning();
    if (!key) { b200seg_set_error("b200seg_set_tuning: key is NULL"); return B200SEG_E_INVALID; }
    if (!strcmp(key, "interleave")) t.interleave = value;
    else if (!strcmp(key, "stats_variant")) t.stats_variant = value;
    else if (!strcmp(key, "emit_path")) t.emit_path = value;
    else if (!strcmp(key, "sort_match")) t.sort_match = value;
    else if (!strcmp(key, "sort_path")) t.sort_path = value;
    else if (!strcmp(key, "dbg")) t.dbg = value;
    else if (!strcmp(key, "pdl")) t.pdl = value;
    else { b200seg_set_error("b200seg_set_tuning: unknown key '%s'", key); return B200SEG_E_INVALID; }
    return 0;
}

extern "C" int b200seg_version(void) { return B200SEG_VERSION; }
extern "C" const char* b200seg_last_error(void) { return g_err; }

// ---- stage events (measurement hook) -------------------------------------------------------------------------
// process-wide on purpose: PyTorch runs backward on its own autograd thread
static cudaEvent_t g_stage_events[B200SEG_N_STAGES] = {nullptr};
static volatile int g_n_stage_events = 0;

extern "C" int b200seg_set_stage_events(void* const* events, int32_t n_events) {
    if (n_events < 0 || n_events > B200SEG_N_STAGES || (n_events > 0 && !events)) {
        b200seg_set_error("b200seg_set_stage_events: need 0 <= n_events <= %d", B200SEG_N_STAGES);
        return B200SEG_E_INVALID;
    }
    g_n_stage_events = n_events;
    for (int i = 0; i < B200SEG_N_STAGES; ++i) g_stage_events[i] = i < n_events ? (cudaEvent_t)events[i] : nullptr;
    return 0;
}

void b200seg_stage(int i, cudaStream_t st) {
    if (i < g_n_stage_events && g_stage_events[i]) cudaEventRecord(g_stage_events[i], st);
}

// ---- confusion-matrix-ready event (data-parallel hook) ---------------------------------------------------------------------
static cudaEvent_t g_cm_event = nullptr;
extern "C" int b200seg_set_confmat_event(void* event) {
    g_cm_event = (cudaEvent_t)event;
    return 0;
}
void b200seg_cm_ready(cudaStream_t st) {
    if (g_cm_event) cudaEventRecord(g_cm_event, st);
}
