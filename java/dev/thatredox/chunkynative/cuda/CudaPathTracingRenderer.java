package dev.thatredox.chunkynative.cuda;

import se.llbit.chunky.renderer.DefaultRenderManager;
import se.llbit.chunky.renderer.Renderer;
import se.llbit.chunky.renderer.ResetReason;
import se.llbit.chunky.renderer.SnapshotControl;
import se.llbit.chunky.renderer.scene.Scene;
import se.llbit.util.TaskTracker;

import java.util.Random;
import java.util.function.BooleanSupplier;

/**
 * Replacement for dev.thatredox.chunkynative.opencl.OpenClPathTracingRenderer: identical Renderer contract
 * (same id "ChunkyClRenderer", so the renderer selector and saved scenes keep working), the JOCL pass loop replaced
 * by ccu_render_passes / ccu_render_merge.  Mirrors chunkyclplugin_b200/renderer.py::CudaPathTracingRenderer, which
 * is the variant exercised by the tests.  NOT COMPILED in the build image (no JDK, no chunky-core jar).
 */
public class CudaPathTracingRenderer implements Renderer {
    private BooleanSupplier postRender = () -> true;
    private final CudaSceneLoader sceneLoader;

    public CudaPathTracingRenderer(CudaSceneLoader sceneLoader) { this.sceneLoader = sceneLoader; }

    @Override public String getId() { return "ChunkyClRenderer"; }
    @Override public String getName() { return "ChunkyClRenderer"; }
    @Override public String getDescription() { return "ChunkyClRenderer"; }
    @Override public void setPostRender(BooleanSupplier callback) { postRender = callback; }
    @Override public boolean autoPostProcess() { return false; }

    @Override
    public void sceneReset(DefaultRenderManager manager, ResetReason reason, int resetCount) {
        sceneLoader.load(resetCount, reason, manager.bufferedScene);
    }

    @Override
    public void render(DefaultRenderManager manager) throws InterruptedException {
        ChunkyCu.Context ctx = sceneLoader.context();
        Scene scene = manager.bufferedScene;
        double[] sampleBuffer = scene.getSampleBuffer();
        sceneLoader.ensureLoad(scene);
        sceneLoader.uploadCamera(scene, null, true);            // ClCamera: settings or pre-generated rays (jittered)
        ctx.renderBegin(scene.width, scene.height);
        try {
            int bufferSppReal = 0;
            int logicalSpp = scene.spp;
            int sceneSpp = scene.spp;
            Random rand = new Random(0);
            SnapshotControl control = manager.getSnapshotControl();
            while (logicalSpp < scene.getTargetSpp()) {
                // all passes up to the next point where the reference would merge: next save event or a full window
                int n = 1024 - bufferSppReal;
                for (int k = 1; k <= n; k++) {
                    int spp = logicalSpp + bufferSppReal + k;
                    if (control.saveSnapshot(scene, spp) || control.saveRenderDump(scene, spp)) { n = k; break; }
                }
                int[] seeds = new int[n];
                for (int i = 0; i < n; i++) seeds[i] = rand.nextInt();
                ctx.renderPasses(seeds);
                bufferSppReal += n;
                scene.spp += n;
                int spp = logicalSpp + bufferSppReal;
                boolean saveEvent = control.saveSnapshot(scene, spp) || control.saveRenderDump(scene, spp);
                if (!scene.shouldFinalizeBuffer() && !saveEvent) {
                    if (postRender.getAsBoolean()) break;
                    if (bufferSppReal < 1024) continue;
                }
                if (postRender.getAsBoolean()) break;
                int passSpp = ctx.renderMerge(sampleBuffer, sceneSpp);
                sceneSpp += passSpp;
                bufferSppReal = 0;
                scene.postProcessFrame(TaskTracker.Task.NONE);
                manager.redrawScreen();
                logicalSpp += passSpp;
            }
        } finally {
            ctx.renderEnd();
        }
    }
}
