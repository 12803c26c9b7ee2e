#ifndef FAKE_CAML_MEMORY_H
#define FAKE_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam3(a, b, c) (void)(a); (void)(b); (void)(c)
#define CAMLparam4(a, b, c, d) (void)(a); (void)(b); (void)(c); (void)(d)
#define CAMLreturn(x) return (x)
#endif
