package dev.thatredox.chunkynative.cuda;

import se.llbit.chunky.Plugin;
import se.llbit.chunky.main.Chunky;
import se.llbit.chunky.model.BlockModel;
import se.llbit.chunky.renderer.postprocessing.PostProcessingFilters;
import se.llbit.log.Log;

/**
 * Replacement for opencl.ChunkyCl (plugin.json "main"): same registrations - path tracing renderer "ChunkyClRenderer",
 * preview renderer "ChunkyClPreviewRenderer", the imposter post-processing filters GAMMA / TONEMAP1 / TONEMAP2 / TONEMAP3
 * (ChunkyCl.java:25-71) - on top of libchunkycu.so instead of JOCL.  The render-controls tab (ui/ChunkyClTab, GpuSelector)
 * stays as it is, with its device list taken from ChunkyCu.deviceCount() / deviceName().  NOT COMPILED in the build image.
 */
public class ChunkyCuPlugin implements Plugin {
    @Override
    public void attach(Chunky chunky) {
        try {
            Class<?> blockModels = BlockModel.class;
        } catch (NoClassDefFoundError e) {
            Log.error("ChunkyCL requires Chunky 2.5.0. Could not load block models.", e);
            return;
        }
        CudaRendererInstance instance;
        try {
            instance = CudaRendererInstance.get();
        } catch (UnsatisfiedLinkError | RuntimeException e) {
            Log.error("Failed to load ChunkyCL. Could not load libchunkycu / no CUDA device.", e);
            return;
        }
        CudaSceneLoader sceneLoader = new CudaSceneLoader(instance.context);
        Chunky.addRenderer(new CudaPathTracingRenderer(sceneLoader));
        Chunky.addPreviewRenderer(new CudaPreviewRenderer(sceneLoader));

        shadow("GAMMA", CudaPostProcessingFilter.Filter.GAMMA, instance);
        shadow("TONEMAP1", CudaPostProcessingFilter.Filter.TONEMAP1, instance);
        shadow("TONEMAP2", CudaPostProcessingFilter.Filter.ACES, instance);
        shadow("TONEMAP3", CudaPostProcessingFilter.Filter.HABLE, instance);
    }

    private static void shadow(String id, CudaPostProcessingFilter.Filter f, CudaRendererInstance instance) {
        PostProcessingFilters.getPostProcessingFilterFromId(id).ifPresent(existing ->
                PostProcessingFilters.addPostProcessingFilter(new CudaPostProcessingFilter(existing, f, instance.context)));
    }
}
