package dev.thatredox.chunkynative.cuda;

import se.llbit.chunky.PersistentSettings;

/**
 * Replacement for opencl.renderer.RendererInstance: the process-wide device context.  The reference enumerates OpenCL
 * platforms / devices, creates a context + queue and JIT-compiles the kernels (RendererInstance.java:31-110); here the
 * device list comes from ccu_device_count / ccu_device_info, the context is one ccu_ctx and the kernels were compiled ahead
 * of time.  The settings key "clDevice" (RendererInstance.java:33, written by ui/GpuSelector.java:72) is kept, so an
 * existing Chunky installation selects the same device index.  NOT COMPILED in the build image (no JDK, no chunky-core jar).
 */
public final class CudaRendererInstance {
    private static CudaRendererInstance instance = null;

    public final int deviceIndex;
    public final ChunkyCu.Context context;

    /** Lazily created; a missing libchunkycu.so / CUDA driver surfaces as UnsatisfiedLinkError as with JOCL (ChunkyCl.java:37-40). */
    public static synchronized CudaRendererInstance get() {
        if (instance == null) instance = new CudaRendererInstance();
        return instance;
    }

    private CudaRendererInstance() {
        int devices = ChunkyCu.deviceCount();
        int preferred = PersistentSettings.settings.getInt("clDevice", 0);
        for (int i = 0; i < devices; i++) System.out.println("  [" + i + "] " + ChunkyCu.deviceName(i));   // RendererInstance.java:67-75
        this.deviceIndex = preferred >= 0 && preferred < devices ? preferred : 0;
        System.out.println("\nUsing device: " + ChunkyCu.deviceName(deviceIndex));
        this.context = new ChunkyCu.Context(deviceIndex);
    }
}
