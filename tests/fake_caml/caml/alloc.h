#ifndef FAKE_CAML_ALLOC_H
#define FAKE_CAML_ALLOC_H
#include "mlvalues.h"
#endif
