// api_jni.cu -- JNI glue (csbwa_jni.inc); compiled in only when a JDK's jni.h is on the include path.
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/csbwa_sw.h"
#include "csbwa_jni.inc"
