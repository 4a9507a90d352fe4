package dev.thatredox.chunkynative.cuda;

import se.llbit.chunky.renderer.postprocessing.PostProcessingFilter;
import se.llbit.chunky.resources.BitmapImage;
import se.llbit.util.TaskTracker;

/**
 * Replacement for opencl.tonemap.GpuPostProcessingFilter / ImposterCombinationGpuPostProcessingFilter: shadows one of
 * Chunky's filters (same name / description / id) and runs it on the device with ccu_tonemap instead of building a
 * cl_program and binding six kernel arguments per frame (GpuPostProcessingFilter.java:40-65).  NOT COMPILED in the build image.
 */
public class CudaPostProcessingFilter implements PostProcessingFilter {
    /** Kernel filter ids, ImposterCombinationGpuPostProcessingFilter.java:11-16. */
    public enum Filter {
        GAMMA(0), TONEMAP1(1), ACES(2), HABLE(3);
        public final int id;
        Filter(int id) { this.id = id; }
    }

    private final PostProcessingFilter imposter;
    private final Filter filter;
    private final ChunkyCu.Context ctx;

    public CudaPostProcessingFilter(PostProcessingFilter imposter, Filter filter, ChunkyCu.Context ctx) {
        this.imposter = imposter;
        this.filter = filter;
        this.ctx = ctx;
    }

    @Override
    public void processFrame(int width, int height, double[] input, BitmapImage output, double exposure, TaskTracker.Task task) {
        ctx.tonemap(width, height, input, output.data, exposure, filter.id);
    }

    @Override public String getName() { return imposter.getName(); }
    @Override public String getDescription() { return imposter.getDescription(); }
    @Override public String getId() { return imposter.getId(); }
}
