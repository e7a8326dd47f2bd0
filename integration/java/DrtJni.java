// NOT COMPILED HERE: no JDK exists in the authoring image. Source of the binding shown in INTEGRATION.md.
final class DrtJni {
  static { System.loadLibrary("drtjni"); }
  static native long create(int device, int cols, int rows, long seed);
  static native int command(long ctx, String line);
  static native int render(long ctx, int accelMode, int[] pixels);   // pixels == rndrdImg.pixels
  static native String lastError(long ctx);
  static native void destroy(long ctx);
}
