/* Stub so the reference unity build does not need CPython headers (mz.h:34 includes <Python.h> only for the Cython bridge). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
