// NOT COMPILED HERE: no JDK exists in the authoring image. Source of the binding shown in INTEGRATION.md.
/* drt_jni.c -- build: gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude drt_jni.c -L. -ldrt -o libdrtjni.so */
#include <jni.h>
#include "drt.h"
#define CTX(h) ((drt_ctx*)(intptr_t)(h))
JNIEXPORT jlong JNICALL Java_rayTracerDistAccelShdPhtnMap_DrtJni_create(JNIEnv* e, jclass c, jint dev, jint cols, jint rows, jlong seed) {
  drt_config cfg = { dev, cols, rows, 0, (uint64_t)seed, 0 }; drt_ctx* ctx = 0;
  return drt_create(&cfg, &ctx) == DRT_OK ? (jlong)(intptr_t)ctx : 0;
}
JNIEXPORT jint JNICALL Java_rayTracerDistAccelShdPhtnMap_DrtJni_command(JNIEnv* e, jclass c, jlong h, jstring line) {
  const char* s = (*e)->GetStringUTFChars(e, line, 0); int rc = drt_scene_command(CTX(h), s); (*e)->ReleaseStringUTFChars(e, line, s); return rc;
}
JNIEXPORT jint JNICALL Java_rayTracerDistAccelShdPhtnMap_DrtJni_render(JNIEnv* e, jclass c, jlong h, jint accel, jintArray pixels) {
  int rc = drt_scene_finalize(CTX(h), accel); if (rc) return rc;
  jint* px = (*e)->GetPrimitiveArrayCritical(e, pixels, 0);            /* PImage.pixels, written in place */
  rc = drt_render(CTX(h), (int32_t*)px, 0);
  (*e)->ReleasePrimitiveArrayCritical(e, pixels, px, 0); return rc;
}
JNIEXPORT jstring JNICALL Java_rayTracerDistAccelShdPhtnMap_DrtJni_lastError(JNIEnv* e, jclass c, jlong h) { return (*e)->NewStringUTF(e, drt_last_error(CTX(h))); }
JNIEXPORT void JNICALL Java_rayTracerDistAccelShdPhtnMap_DrtJni_destroy(JNIEnv* e, jclass c, jlong h) { drt_destroy(CTX(h)); }
