/* time_shim.c -- LD_PRELOAD shim so the reference's generators, which seed with
 * srand(time(NULL)) (src/write_data.c:16, src/write_query.c:18), become reproducible
 * without touching their source:  HVS_SEED=7 LD_PRELOAD=oracle/_ref/time_shim.so write_data f N
 * TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
#include <time.h>
time_t time(time_t *t)
{
    const char *s = getenv("HVS_SEED");
    time_t v = s ? (time_t)atol(s) : (time_t)1;
    if (t) *t = v;
    return v;
}
