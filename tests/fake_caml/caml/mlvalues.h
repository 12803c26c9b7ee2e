/* A minimal stand-in for OCaml's <caml/mlvalues.h>, only to compile and drive integration/pmn_stubs.c in the tests
 * (this image has no OCaml toolchain).  Strings are plain C strings behind the `value` word. */
#ifndef FAKE_CAML_MLVALUES_H
#define FAKE_CAML_MLVALUES_H
#include <stdint.h>
typedef intptr_t value;
#define Val_unit ((value)1)
#define String_val(v) ((const char *)(v))
#define CAMLprim
#endif
