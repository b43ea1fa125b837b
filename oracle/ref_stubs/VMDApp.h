/* stand-in for VMDApp.h + WKFThreads.h (not in the snapshot).  oracle/_ref build only.
   WKF calls used by MDFF.C:58,65,89,96,118,136,140,141; VMDApp calls by Voltool.C:217-223. */
#ifndef STUB_VMDAPP_H
#define STUB_VMDAPP_H
#include <pthread.h>
#include <string.h>
typedef pthread_mutex_t wkf_mutex_t;
typedef struct wkf_tasktile_struct { int start; int end; } wkf_tasktile_t;
#define WKF_SCHED_DONE -1
#define WKF_SCHED_CONTINUE 0
extern "C" {
int wkf_thread_numprocessors(void);
int wkf_mutex_init(wkf_mutex_t *);
int wkf_mutex_lock(wkf_mutex_t *);
int wkf_mutex_unlock(wkf_mutex_t *);
int wkf_mutex_destroy(wkf_mutex_t *);
int wkf_threadlaunch(int numprocs, void *clientdata, void *fctn(void *), wkf_tasktile_t *tile);
int wkf_threadlaunch_getdata(void *thrparms, void **clientdata);
int wkf_threadlaunch_next_tile(void *thrparms, int reqsize, wkf_tasktile_t *tile);
}
class VMDApp {
public:
  int molecule_new(const char *, int, int);
  int molecule_add_volumetric(int, const char *, const double *, const double *, const double *, const double *,
                              int, int, int, float *);
  int molecule_set_style(const char *);
  int molecule_addrep(int);
};
#endif
