// NOT COMPILED HERE: no JDK exists in the authoring image. Source of the binding shown in INTEGRATION.md.
// DrtNative.java -- drop into src/rayTracerDistAccelShdPhtnMap/
package rayTracerDistAccelShdPhtnMap;
import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

final class DrtNative implements AutoCloseable {
  private static final Linker L = Linker.nativeLinker();
  private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("drt.lib", "libdrt.so"), Arena.global());
  private static MethodHandle h(String n, FunctionDescriptor d) { return L.downcallHandle(LIB.find(n).orElseThrow(), d); }
  // struct drt_config { int32 device, cols, rows, counters; uint64 seed; int64 batch_rays; }   (include/drt.h)
  private static final StructLayout CONFIG = MemoryLayout.structLayout(JAVA_INT.withName("device"), JAVA_INT.withName("cols"),
      JAVA_INT.withName("rows"), JAVA_INT.withName("counters"), JAVA_LONG.withName("seed"), JAVA_LONG.withName("batch_rays"));
  private static final MethodHandle CREATE   = h("drt_create",         FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle DESTROY  = h("drt_destroy",        FunctionDescriptor.ofVoid(ADDRESS));
  private static final MethodHandle ERR      = h("drt_last_error",     FunctionDescriptor.of(ADDRESS, ADDRESS));
  private static final MethodHandle RESET    = h("drt_scene_reset",    FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle COMMAND  = h("drt_scene_command",  FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle TEXDIR   = h("drt_set_texture_dir",FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle FINALIZE = h("drt_scene_finalize", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  private static final MethodHandle RENDER   = h("drt_render",         FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
  // multi-GPU inside the library (include/drt.h: drt_comm_unique_id / drt_comm_init / drt_render_distributed), one DrtNative per GPU
  private static final MethodHandle COMM_ID   = h("drt_comm_unique_id",     FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle COMM_INIT = h("drt_comm_init",          FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
  private static final MethodHandle RENDER_D  = h("drt_render_distributed", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));

  private final Arena arena = Arena.ofConfined();
  private final MemorySegment ctx;
  final int cols, rows;

  DrtNative(int device, int cols, int rows, long seed) throws Throwable {
    this.cols = cols; this.rows = rows;
    MemorySegment cfg = arena.allocate(CONFIG);
    cfg.set(JAVA_INT, 0, device); cfg.set(JAVA_INT, 4, cols); cfg.set(JAVA_INT, 8, rows); cfg.set(JAVA_INT, 12, 0);
    cfg.set(JAVA_LONG, 16, seed); cfg.set(JAVA_LONG, 24, 0L);
    MemorySegment out = arena.allocate(ADDRESS);
    check((int) CREATE.invoke(cfg, out));          // DRT_ERR_NO_DEVICE (-1) when there is no GPU: there is no CPU fallback
    ctx = out.get(ADDRESS, 0);
  }
  private void check(int rc) throws Throwable {
    if (rc != 0) throw new IllegalStateException("drt error " + rc + ": " + ((MemorySegment) ERR.invoke(ctx)).reinterpret(4096).getString(0));
  }
  void reset() throws Throwable { check((int) RESET.invoke(ctx)); }
  void command(String line) throws Throwable { try (Arena a = Arena.ofConfined()) { check((int) COMMAND.invoke(ctx, a.allocateFrom(line))); } }
  void textureDir(String dir) throws Throwable { try (Arena a = Arena.ofConfined()) { check((int) TEXDIR.invoke(ctx, a.allocateFrom(dir))); } }
  /** finalize + render; returns PImage.pixels-compatible ARGB ints (row-major, alpha 0xFF). */
  int[] render(int accelMode) throws Throwable {
    check((int) FINALIZE.invoke(ctx, accelMode));
    try (Arena a = Arena.ofConfined()) {
      MemorySegment px = a.allocate(JAVA_INT, (long) cols * rows);
      check((int) RENDER.invoke(ctx, px, MemorySegment.NULL));
      return px.toArray(JAVA_INT);
    }
  }
  /** rank 0: the 128-byte communicator id the host ships to the other ranks by any transport */
  static byte[] commUniqueId() throws Throwable { try (Arena a = Arena.ofConfined()) { MemorySegment id = a.allocate(128); int rc = (int) COMM_ID.invoke(id); if (rc != 0) throw new IllegalStateException("drt_comm_unique_id " + rc); return id.toArray(JAVA_BYTE); } }
  void commInit(byte[] id128, int world, int rank) throws Throwable { try (Arena a = Arena.ofConfined()) { check((int) COMM_INIT.invoke(ctx, a.allocateFrom(JAVA_BYTE, id128), world, rank)); } }
  /** collective: every rank renders its interleaved 8-row chunks; rank 0 gets the assembled frame, the others null */
  int[] renderDistributed(int accelMode, int rank) throws Throwable {
    check((int) FINALIZE.invoke(ctx, accelMode));
    try (Arena a = Arena.ofConfined()) {
      MemorySegment px = rank == 0 ? a.allocate(JAVA_INT, (long) cols * rows) : MemorySegment.NULL;
      check((int) RENDER_D.invoke(ctx, px, MemorySegment.NULL, 8, 1, MemorySegment.NULL));
      return rank == 0 ? px.toArray(JAVA_INT) : null;
    }
  }
  @Override public void close() { try { DESTROY.invoke(ctx); } catch (Throwable t) { } arena.close(); }
}
