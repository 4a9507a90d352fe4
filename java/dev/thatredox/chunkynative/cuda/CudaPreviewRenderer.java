package dev.thatredox.chunkynative.cuda;

import se.llbit.chunky.renderer.DefaultRenderManager;
import se.llbit.chunky.renderer.Renderer;
import se.llbit.chunky.renderer.ResetReason;
import se.llbit.chunky.renderer.scene.Scene;

import java.util.function.BooleanSupplier;

/**
 * Replacement for opencl.OpenClPreviewRenderer (same id "ChunkyClPreviewRenderer"): one ccu_preview call instead of 18
 * clSetKernelArg + NDRange + blocking read (OpenClPreviewRenderer.java:47-115).  NOT COMPILED in the build image.
 */
public class CudaPreviewRenderer implements Renderer {
    private BooleanSupplier postRender = () -> true;
    private final CudaSceneLoader sceneLoader;

    public CudaPreviewRenderer(CudaSceneLoader sceneLoader) { this.sceneLoader = sceneLoader; }

    @Override public String getId() { return "ChunkyClPreviewRenderer"; }
    @Override public String getName() { return "ChunkyClPreviewRenderer"; }
    @Override public String getDescription() { return "ChunkyClPreviewRenderer"; }
    @Override public void setPostRender(BooleanSupplier callback) { postRender = callback; }
    @Override public boolean autoPostProcess() { return false; }

    @Override
    public void sceneReset(DefaultRenderManager manager, ResetReason reason, int resetCount) {
        sceneLoader.load(resetCount, reason, manager.bufferedScene);
    }

    @Override
    public void render(DefaultRenderManager manager) throws InterruptedException {
        Scene scene = manager.bufferedScene;
        ChunkyCu.Context ctx = sceneLoader.context();
        sceneLoader.ensureLoad(scene);
        sceneLoader.uploadCamera(scene, null, false);             // camera.generate(null, false), :68
        ctx.renderBegin(scene.width, scene.height);
        try {
            ctx.preview(scene.getBackBuffer().data);               // ARGB ints, one per pixel
        } finally {
            ctx.renderEnd();
        }
        manager.redrawScreen();
        postRender.getAsBoolean();
    }
}
