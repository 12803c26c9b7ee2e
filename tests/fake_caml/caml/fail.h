#ifndef FAKE_CAML_FAIL_H
#define FAKE_CAML_FAIL_H
/* raises Failure: never returns (the driver catches it with setjmp) */
void caml_failwith(const char *msg) __attribute__((noreturn));
#endif
