package dev.thatredox.chunkynative.cuda;

import se.llbit.chunky.renderer.DefaultRenderManager;
import se.llbit.chunky.renderer.Renderer;
import se.llbit.chunky.renderer.ResetReason;
import se.llbit.chunky.renderer.SnapshotControl;
import se.llbit.chunky.renderer.scene.Scene;
import se.llbit.util.TaskTracker;

import se.llbit.chunky.main.Chunky;

import java.util.Random;
import java.util.concurrent.ForkJoinTask;
import java.util.concurrent.locks.ReentrantLock;
import java.util.function.BooleanSupplier;

/**
 * Replacement for dev.thatredox.chunkynative.opencl.OpenClPathTracingRenderer: identical Renderer contract
 * (same id "ChunkyClRenderer", so the renderer selector and saved scenes keep working), the JOCL pass loop replaced
 * by ccu_render_passes / ccu_render_window_close + ccu_render_window_merge.  Mirrors chunkyclplugin_b200/renderer.py::
 * CudaPathTracingRenderer, which is the variant exercised by the tests.  NOT COMPILED in the build image (no JDK, no
 * chunky-core jar).
 *
 * Threading follows the reference (OpenClPathTracingRenderer.java:56,97-98,146-151,172-182): the pass loop runs on Chunky's
 * render thread under renderLock; the camera-ray regeneration and the buffer merge run as Chunky.getCommonThreads() tasks.
 * One C-ABI call covers as many passes as fit in about 100 ms of device time (the reference launches one pass at a time and
 * polls postRender at most every 100 ms, :153-157), never more than up to the next merge point.
 */
public class CudaPathTracingRenderer implements Renderer {
    private static final int MERGE_WINDOW = 1024;            // :158
    private static final double CALLBACK_MS = 100.0;         // :153-157
    private static final int CAMERA_REGEN_PASSES = 8;        // generated-ray cameras: passes per ray set at most

    private BooleanSupplier postRender = () -> true;
    private final CudaSceneLoader sceneLoader;
    private double msPerPass = -1;                           // device time per pass of the last batch, kept across renders

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

    private int batchLimit() {
        if (msPerPass < 0) return 1;                          // first call: one pass, to learn what a pass costs
        return Math.max(1, (int) (CALLBACK_MS / Math.max(msPerPass, 1e-3)));
    }

    @Override
    public void render(DefaultRenderManager manager) throws InterruptedException {
        ChunkyCu.Context ctx = sceneLoader.context();
        ReentrantLock renderLock = new ReentrantLock();
        Scene scene = manager.bufferedScene;
        double[] sampleBuffer = scene.getSampleBuffer();
        sceneLoader.ensureLoad(scene);
        final boolean needGenerate = sceneLoader.uploadCamera(scene, renderLock, true);   // ClCamera: settings, or jittered pre-generated rays
        ctx.renderBegin(scene.width, scene.height);
        ForkJoinTask<?> cameraGenTask = Chunky.getCommonThreads().submit(() -> 0);       // :97-98
        ForkJoinTask<?> bufferMergeTask = Chunky.getCommonThreads().submit(() -> 0);
        try {
            int bufferSppReal = 0;
            int logicalSpp = scene.spp;
            final int[] sceneSpp = {scene.spp};
            long lastCallback = 0;
            Random rand = new Random(0);
            SnapshotControl control = manager.getSnapshotControl();
            while (logicalSpp < scene.getTargetSpp()) {
                // all passes up to the next point where the reference would merge (next save event or a full window), capped
                int n = Math.max(1, Math.min(MERGE_WINDOW - bufferSppReal, batchLimit()));
                if (needGenerate) n = Math.min(n, CAMERA_REGEN_PASSES);
                for (int k = 1; k <= n; k++) {
                    int spp = logicalSpp + bufferSppReal + k;
                    if (control.saveSnapshot(scene, spp) || control.saveRenderDump(scene, spp)) { n = k; break; }
                }
                int[] seeds = new int[n];
                for (int i = 0; i < n; i++) seeds[i] = rand.nextInt();                   // :106-107
                renderLock.lock();
                try {
                    ctx.renderPasses(seeds);                                              // :108-141
                } finally {
                    renderLock.unlock();
                }
                msPerPass = ctx.lastKernelMs() / n;
                bufferSppReal += n;
                scene.spp += n;
                if (needGenerate && cameraGenTask.isDone()) {                             // :146-148
                    cameraGenTask = Chunky.getCommonThreads().submit(() -> { sceneLoader.uploadCamera(scene, renderLock, true); });
                }
                int spp = logicalSpp + bufferSppReal;
                boolean saveEvent = control.saveSnapshot(scene, spp) || control.saveRenderDump(scene, spp);
                if (bufferMergeTask.isDone() || saveEvent) {                              // :150
                    if (!scene.shouldFinalizeBuffer() && !saveEvent) {
                        long time = System.currentTimeMillis();
                        if (time - lastCallback > CALLBACK_MS && !manager.shouldFinalize()) {
                            lastCallback = time;
                            if (postRender.getAsBoolean()) break;
                        }
                        if (bufferSppReal < MERGE_WINDOW) continue;
                    }
                    bufferMergeTask.join();
                    if (postRender.getAsBoolean()) break;
                    // :164-173.  The window is closed on this thread (the next passes go to the library's second buffer, the
                    // read-back starts); the merge itself runs as a common-pool task like the reference's.
                    final int sampSpp = sceneSpp[0];
                    final int passSpp = ctx.windowClose();
                    bufferSppReal = 0;
                    bufferMergeTask = Chunky.getCommonThreads().submit(() -> {
                        ctx.windowMerge(sampleBuffer, sampSpp);
                        sceneSpp[0] += passSpp;
                        scene.postProcessFrame(TaskTracker.Task.NONE);
                        manager.redrawScreen();
                    });
                    logicalSpp += passSpp;
                    if (saveEvent) {
                        bufferMergeTask.join();
                        if (postRender.getAsBoolean()) break;
                    }
                }
            }
        } finally {
            cameraGenTask.join();
            bufferMergeTask.join();
            ctx.renderEnd();
        }
    }
}
