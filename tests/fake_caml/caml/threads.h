#ifndef FAKE_CAML_THREADS_H
#define FAKE_CAML_THREADS_H
void caml_release_runtime_system(void);
void caml_acquire_runtime_system(void);
#endif
