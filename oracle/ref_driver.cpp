// oracle/ref_driver.cpp -- glue that runs the REFERENCE's own host code (CudaTracerLib, compiled by oracle/build_ref.sh
// from /root/reference) on the flat scene view of include/ctl_b200.h.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header): loaded by tests/ref_binding.py to validate the CPU restatement
// and by bench.py's cpu_baseline / --impl reference legs.  The product never links or loads it.
//
// What is the reference's code here: TracerayTemplate (BVH traversal), PathTrace<DIRECT> (Integrators/PathTracer.cu:10-113),
// UniformSampleOneLight/EstimateDirect, TraceResult::getBsdfSample, TriangleData::fillDG, BSDFALL (diffuse, roughconductor,
// dielectric), MicrofacetDistribution, DiffuseLight, ShapeSet::SamplePosition, PerspectiveSensor::sampleRayDifferential,
// SequenceSampler + SamplingSequenceGeneratorHost + CudaRNG (XORWOW), Image::AddSample.
// traceRay / __traceRay_internal__ / fillDG are the reference's own lines too (Kernel/TraceHelper.cu:62-180, 274-307, cut out
// of the file by oracle/build_ref.sh because the rest of it is CUDA-12-removed texture<> declarations and kernels).
// What is written here: the scene/sampler globals (TraceHelper.cu:23-32), CUDA-runtime stubs backed by host memory, and
// the packing of ctl_scene_view into KernelDynamicScene.
#include <Kernel/TraceHelper.h>
#include <Kernel/TraceAlgorithms.h>
#include <Kernel/Sampler.h>
#include <Engine/Mesh.h>
#include <Engine/TriangleData.h>
#include <Engine/Material.h>
#include <Engine/TriIntersectorData.h>
#include <Engine/ShapeSet.h>
#include <Engine/Image.h>
#include <SceneTypes/Node.h>
#include <SceneTypes/Light.h>
#include <SceneTypes/Sensor.h>
#include <SceneTypes/BSDF.h>
#include <Engine/SpatialStructures/BVH/BVHTraversal.h>
#include <Base/CudaMemoryManager.h>
#include <Base/Platform.h>
#include <thread>
#include <vector>
#include <atomic>
#include <mutex>
#include <cstring>
#include <cstdlib>
#include <cstddef>
#include "ctl_b200.h"

// ---- CUDA runtime entry points used by host-side buffers: plain host memory (no GPU involved) -----------------
extern "C" {
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memcpy(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemset(void* dst, int v, size_t n) { memset(dst, v, n); return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "cuda stub"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
}

namespace CudaTracerLib {

std::map<void*, CudaMemoryEntry> CudaMemoryManager::alloced_entries;
std::vector<CudaMemoryEntry> CudaMemoryManager::freed_entries;
cudaError_t CudaMemoryManager::Cuda_malloc_managed(void** v, size_t i, const std::string&) { *v = malloc(i ? i : 1); return cudaSuccess; }
cudaError_t CudaMemoryManager::Cuda_free_managed(void* v, const std::string&) { free(v); return cudaSuccess; }

// ---- globals of Kernel/TraceHelper.cu:23-32 --------------------------------------------------------------------
KernelDynamicScene g_SceneDataHost;
unsigned int g_RayTracedCounterHost;
CudaStaticWrapper<SamplerData> g_SamplerDataHost;

// ---- traceRay / fillDG: the reference's own text (Kernel/TraceHelper.cu:62-180, 274-307), extracted by oracle/build_ref.sh
#include <Kernel/TraceHelper_host.inc>

} // namespace CudaTracerLib

// ---- stubs for subsystems off the hot path that the compiled TUs reference ---------------------------
#include "ref_stubs.inc"

#include <Integrators/PathTracer_host.inc>   // the reference's PathTrace<DIRECT> (Integrators/PathTracer.cu:1-170)
#include <SceneTypes/Filter.h>
#include <SceneTypes/Dispersion.h>
#include <Engine/Mesh.h>
#include <Engine/MeshLoader/MeshCompiler.h>
#include <Base/FileStream.h>
#include <Kernel/PixelVarianceBuffer.h>
#include <Kernel/ImagePipeline/PostProcess/ToneMapPostProcess.h>
namespace CudaTracerLib {
#include <Kernel/ImagePipeline/Filter/evalFilter_host.inc>   // the reference's evalFilter (CanonicalFilter.cu:6-27)
}

// ---- WavefrontPathTracer: the reference's own pathIterateKernel<NEE> (Integrators/PseudoRealtime/WavefrontPathTracer.cu:51-164)
// over the reference's own DoubleRayBuffer (Kernel/DoubleRayBuffer.h, unpatched), run by one host thread = the serial schedule of its atomics.
namespace CudaTracerLib {
static inline unsigned int atomicInc(unsigned int* a, unsigned int limit) { unsigned int o = *a; *a = (o >= limit) ? 0u : o + 1u; return o; } // CUDA semantics, one thread
static int g_wpt_threads = 1;
static std::vector<int> g_wpt_calls;   // N of every __internal__IntersectBuffers call since the last clear
// __internal__IntersectBuffers (Kernel/TraceHelper.cu:736-746) on the host: the reference's traceRay (t in (rayEps, FLT_MAX), what the queue's
// rays carry: DoubleRayBuffer.h:236-237) + the reference's traversalResult::fromResult (16-bit barycentrics); a miss is (FLT_MAX, -1, -1, 0)
// like the kernel's store (TraceHelper.cu:722-731).  Closest hit only (WavefrontPathTracer never asks for any-hit).
void __internal__IntersectBuffers(int N, traversalRay* rays, traversalResult* res, bool, bool)
{
	g_wpt_calls.push_back(N);
	std::atomic<int> next(0);
	auto work = [&]() {
		for (;;) {
			int i0 = next.fetch_add(256); if (i0 >= N) break;
			for (int i = i0; i < N && i < i0 + 256; i++) {
				TraceResult r2 = traceRay(Ray(rays[i].a.getXYZ(), rays[i].b.getXYZ()));
				if (r2.hasHit()) res[i].fromResult(&r2, g_SceneDataHost);
				else { res[i].dist = r2.m_fDist; res[i].nodeIdx = -1; res[i].triIdx = -1; res[i].bCoords = 0; }
			}
		}
	};
	std::vector<std::thread> th;
	for (int t = 1; t < g_wpt_threads; t++) th.emplace_back(work);
	work();
	for (auto& t : th) t.join();
}
}
#include <Kernel/DoubleRayBuffer.h>
#include <Math/half.h>
#include <Math/Compression.h>
namespace CudaTracerLib {
#include <Integrators/PseudoRealtime/WavefrontPT_payload_host.inc>   // struct WavefrontPTRayData, WavefrontPathTracerBuffer (WavefrontPathTracer.h:11-24)
CudaStaticWrapper<WavefrontPathTracerBuffer> g_ray_buffer;             // WavefrontPathTracer.cu:14-15
DeviceDepthImage g_DepthImageWPT;                                      // Kernel/Tracer.h:16-32; never stored to (depthImage = false)
#include <Integrators/PseudoRealtime/WavefrontPT_iterate_host.inc>   // the reference's pathIterateKernel<NEXT_EVENT_EST> (WavefrontPathTracer.cu:51-164)
}

// NonLocalMeansFilter kernels as host functions (build_ref.sh note 13)
#include <Kernel/ImagePipeline/Filter/NonLocalMeansFilter.h>
namespace CudaTracerLib {
namespace nlm_host {
struct U3 { unsigned x, y, z; };
static U3 h_threadIdx, h_blockIdx, h_blockDim;
static inline void h_syncthreads() {}
#define threadIdx nlm_host::h_threadIdx
#define blockIdx nlm_host::h_blockIdx
#define blockDim nlm_host::h_blockDim
#define __syncthreads nlm_host::h_syncthreads
}
#include <Kernel/ImagePipeline/Filter/NonLocalMeans_host.inc>
template <typename K> static void nlm_launch(unsigned gx, unsigned gy, K kernel) {
	nlm_host::h_blockDim = {16, 16, 1};
	for (unsigned by = 0; by < gy; by++) for (unsigned bx = 0; bx < gx; bx++) {
		nlm_host::h_blockIdx = {bx, by, 0};
		for (int sweep = 0; sweep < 2; sweep++)
			for (unsigned ty = 0; ty < 16; ty++) for (unsigned tx = 0; tx < 16; tx++) { nlm_host::h_threadIdx = {tx, ty, 0}; kernel(); }
	}
}
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef __syncthreads
} // namespace CudaTracerLib

using namespace CudaTracerLib;

namespace {

struct ShapeSetMirror { unsigned int areaIdx, areaLen, triIdx, triLen; float sumArea; unsigned int count; }; // Engine/ShapeSet.h:47-52
static_assert(sizeof(ShapeSetMirror) == sizeof(ShapeSet), "ShapeSet layout");
static_assert(sizeof(BVHNodeData) == sizeof(ctl_bvh_node) && sizeof(TriIntersectorData) == sizeof(ctl_woop_tri) && sizeof(TriangleData) == sizeof(ctl_tri_data) &&
              sizeof(KernelMesh) == sizeof(ctl_mesh) && sizeof(Node) == sizeof(ctl_node) && sizeof(ShapeSet::triData) == sizeof(ctl_light_tri) &&
              sizeof(PixelData) == sizeof(ctl_pixel_data) && sizeof(float4x4) == 64, "data-surface layouts must be byte-identical");
static_assert(offsetof(Node, m_uLights) == offsetof(ctl_node, n_lights) && offsetof(KernelMesh, m_uStdMaterialOffset) == offsetof(ctl_mesh, mat_offset) &&
              offsetof(ShapeSet::triData, area) == offsetof(ctl_light_tri, area) && offsetof(PixelData, weightSum) == offsetof(ctl_pixel_data, weight_sum), "field offsets");

struct RefScene {
	std::vector<Material> mats; std::vector<Light> lights; std::vector<char> anim; std::vector<float> lightPdf;
};
RefScene* g_scene = nullptr;
SamplingSequenceGeneratorHost<IndependantSamplingSequenceGenerator>* g_gen = nullptr;
unsigned g_passes_generated = 0;
std::mutex g_mutex;

void pack_scene(const ctl_scene_view& v)
{
	delete g_scene; g_scene = new RefScene();
	RefScene& R = *g_scene;
	KernelDynamicScene& K = g_SceneDataHost;
	memset((void*)&K, 0, sizeof(K));
	K.m_sTriData = {(TriangleData*)v.tri_data, v.n_tri_data, v.n_tri_data};
	K.m_sBVHIntData = {(TriIntersectorData*)v.woop, v.n_woop, v.n_woop};
	K.m_sBVHNodeData = {(BVHNodeData*)v.bvh_nodes, v.n_bvh_nodes, v.n_bvh_nodes};
	K.m_sBVHIndexData = {(TriIntersectorData2*)v.tri_index, v.n_tri_index, v.n_tri_index};
	K.m_sMeshData = {(KernelMesh*)v.meshes, v.n_meshes, v.n_meshes};
	K.m_sNodeData = {(Node*)v.nodes, v.n_nodes, v.n_nodes};
	K.m_sSceneBVH.m_sStartNode = v.scene_start_node; K.m_sSceneBVH.m_uNumNodes = v.n_scene_bvh_nodes;
	K.m_sSceneBVH.m_pNodes = (BVHNodeData*)v.scene_bvh_nodes;
	K.m_sSceneBVH.m_pNodeTransforms = (float4x4*)v.node_xf; K.m_sSceneBVH.m_pInvNodeTransforms = (float4x4*)v.node_inv_xf;
	K.m_uEnvMapIndex = UINT_MAX;
	K.m_sBox = AABB(Vec3f(v.box_min[0], v.box_min[1], v.box_min[2]), Vec3f(v.box_max[0], v.box_max[1], v.box_max[2]));
	K.doAlphaMapping = false;
	K.m_rayTraceEps = v.ray_eps;
	// materials: ctl_material -> Material with the reference's own BSDF objects
	R.mats.resize(v.n_materials);
	for (unsigned i = 0; i < v.n_materials; i++) {
		const ctl_material& m = v.materials[i];
		Material M("m");
		M.NodeLightIndex = m.node_light_index;
		Spectrum refl(m.reflectance[0], m.reflectance[1], m.reflectance[2]);
		if (m.bsdf_type == CTL_BSDF_DIFFUSE) M.bsdf.SetData(diffuse(CreateTexture(refl)));
		else if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR)
			M.bsdf.SetData(roughconductor(m.distr_type == CTL_DISTR_GGX ? MicrofacetDistribution::EGGX : MicrofacetDistribution::EBeckmann,
			                              Spectrum(m.eta[0], m.eta[1], m.eta[2]), Spectrum(m.k[0], m.k[1], m.k[2]),
			                              CreateTexture(Spectrum(m.alpha_u)), CreateTexture(Spectrum(m.alpha_v)), CreateTexture(refl)));
		else M.bsdf.SetData(dielectric(m.eta[0], refl, Spectrum(m.transmittance)));
		M.bsdf.As()->m_enableTwoSided = (m.flags & CTL_MAT_TWO_SIDED) != 0;
		R.mats[i] = M;
	}
	K.m_sMatData = {R.mats.data(), v.n_materials, v.n_materials};
	// m_sAnimData: per light [area CDF (count+1 floats) | triData[count]] as ShapeSet expects (Engine/ShapeSet.cu:31-34)
	R.lights.resize(v.n_lights_buf);
	size_t total = 0;
	for (unsigned i = 0; i < v.n_lights_buf; i++) total += ((v.lights[i].count + 1) * 4 + 15) / 16 * 16 + (size_t)v.lights[i].count * 64;
	R.anim.assign(total + 64, 0);
	char* base = (char*)(((uintptr_t)R.anim.data() + 15) & ~(uintptr_t)15);
	size_t off = 0;
	for (unsigned i = 0; i < v.n_lights_buf; i++) {
		const ctl_light& L = v.lights[i];
		ShapeSetMirror sm; sm.areaIdx = (unsigned)off; sm.areaLen = (L.count + 1) * 4;
		memcpy(base + off, v.light_cdf_data + L.cdf_offset, sm.areaLen); off += (sm.areaLen + 15) / 16 * 16;
		sm.triIdx = (unsigned)off; sm.triLen = L.count * 64;
		memcpy(base + off, v.light_tris + L.tri_offset, sm.triLen); off += sm.triLen;
		sm.sumArea = L.sum_area; sm.count = L.count;
		ShapeSet ss; memcpy((void*)&ss, &sm, sizeof(sm));
		Light l; l.SetData(DiffuseLight(Spectrum(L.radiance[0], L.radiance[1], L.radiance[2]), ss, L.node_idx));
		R.lights[i] = l;
	}
	K.m_sAnimData = {base, (unsigned)off, (unsigned)off};
	K.m_sLightBuf = {R.lights.data(), v.n_lights_buf, v.n_lights_buf};
	K.m_numLights = v.num_lights;
	R.lightPdf.assign(v.n_lights_buf ? v.n_lights_buf : 1, 0.0f);
	for (unsigned i = 0; i < v.num_lights; i++) {
		K.m_pLightIndices[i] = v.light_indices[i]; K.m_pLightCDF[i] = v.light_cdf[i];
		R.lightPdf[v.light_indices[i]] = v.light_cdf[i] - (i ? v.light_cdf[i - 1] : 0.0f);
	}
	K.m_pLightPDF = R.lightPdf.data();
	// camera: PerspectiveSensor with the matrices of the view (SceneTypes/Sensor.cu:76-96)
	PerspectiveSensor ps((int)v.camera.resolution[0], (int)v.camera.resolution[1], 60.0f);
	memcpy((void*)&ps.m_sampleToCamera, v.camera.sample_to_camera, 64);
	ps.m_cameraToSample = ps.m_sampleToCamera.inverse();
	memcpy((void*)&ps.toWorld, v.camera.to_world, 64);
	ps.m_invResolution = Vec2f(v.camera.inv_resolution[0], v.camera.inv_resolution[1]);
	ps.m_dx = ps.m_sampleToCamera.TransformPoint(Vec3f(ps.m_invResolution.x, 0.0f, 0.0f)) - ps.m_sampleToCamera.TransformPoint(Vec3f(0.0f));
	ps.m_dy = ps.m_sampleToCamera.TransformPoint(Vec3f(0.0f, ps.m_invResolution.y, 0.0f)) - ps.m_sampleToCamera.TransformPoint(Vec3f(0.0f));
	K.m_Camera.SetData(ps);
}

} // namespace

extern "C" {

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// Renders passes [pass_first, pass_first + n_passes) of the window into img (PixelData[w*h], accumulated) with the reference's
// PathTrace<DIRECT>; body of pathKernel2 (Integrators/PathTracer.cu:184-193) looped over the pixels.  Returns the ray count
// (g_RayTracedCounterHost, every traceRay call).
unsigned long long ref_render(const ctl_scene_view* view, int w, int h, int x0, int y0, int x1, int y1, int pass_first, int n_passes,
                              int max_path_length, int rr_start, int direct, ctl_pixel_data* img, int n_threads)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	pack_scene(*view);
	const bool regularization = (direct & 2) != 0;   // bit 1 of `direct`: KEY_Regularization -> the reference's PathTraceRegularization<DIRECT> (Integrators/PathTracer.cu:115-170)
	direct &= 1;
	static bool sampler_init = false;
	if (!sampler_init) { new (&(*g_SamplerDataHost)) SamplerData(4096, 30); sampler_init = true; } // Kernel/TraceHelper.cu:257
	// fresh tracer: sample stream restarts; batches before pass_first are generated and discarded
	delete g_gen; g_gen = new SamplingSequenceGeneratorHost<IndependantSamplingSequenceGenerator>();
	for (int p = 0; p < pass_first; p++) g_gen->Compute(*g_SamplerDataHost);
	g_RayTracedCounterHost = 0;
	Image I(w, h);
	I.Clear();
	for (int yy = 0; yy < h; yy++) for (int xx = 0; xx < w; xx++) memcpy((void*)&I.getPixelData(xx, yy), &img[yy * w + xx], sizeof(PixelData));
	if (n_threads < 1) n_threads = 1;
	struct Smp { float x, y; Spectrum L; };
	for (int p = 0; p < n_passes; p++) {
		g_gen->Compute(*g_SamplerDataHost);
		std::vector<std::vector<Smp>> per_row((size_t)(y1 - y0));
		std::atomic<int> next(y0);
		auto work = [&]() {
			for (;;) {
				int y = next.fetch_add(1);
				if (y >= y1) break;
				std::vector<Smp>& out = per_row[(size_t)(y - y0)];
				out.reserve((size_t)(x1 - x0));
				for (int x = x0; x < x1; x++) {
					auto rng = g_SamplerData((unsigned)(y * w + x));
					NormalizedT<Ray> r, rX, rY;
					Vec2f pX = Vec2f((float)x, (float)y) + rng.randomFloat2();
					Spectrum imp = g_SceneData.sampleSensorRay(r, rX, rY, pX, rng.randomFloat2());
					Spectrum col;
					if (regularization) { // RenderBlock's mollifier radius (Integrators/PathTracer.cu:196-199); only read for lights this path does not have
						AABB box = g_SceneData.m_sBox;
						float r0 = (box.maxV - box.minV).sum() / 100, m = math::pow(math::pow(r0, float(2)) / math::pow(float(pass_first + p + 1), 0.5f * (1 - 0.75f)), 1.0f / 2.0f);
						col = imp * (direct ? PathTraceRegularization<true>(r, rX, rY, rng, m, max_path_length, rr_start) : PathTraceRegularization<false>(r, rX, rY, rng, m, max_path_length, rr_start));
					} else
						col = imp * (direct ? PathTrace<true>(r, rX, rY, rng, max_path_length, rr_start) : PathTrace<false>(r, rX, rY, rng, max_path_length, rr_start));
					out.push_back({pX.x, pX.y, col});
				}
			}
		};
		std::vector<std::thread> th;
		for (int t = 1; t < n_threads; t++) th.emplace_back(work);
		work();
		for (auto& t : th) t.join();
		for (auto& row : per_row) for (auto& s : row) I.AddSample(s.x, s.y, s.L); // the reference's Image::AddSample, in pixel order
	}
	for (int yy = 0; yy < h; yy++) for (int xx = 0; xx < w; xx++) memcpy(&img[yy * w + xx], (void*)&I.getPixelData(xx, yy), sizeof(PixelData));
	I.Free();
	return g_RayTracedCounterHost;
}

// WavefrontPathTracer::DoRender (Integrators/PseudoRealtime/WavefrontPathTracer.cu:166-191) for passes [pass_first, pass_first + n_passes) of a
// fresh tracer, one host thread for the queue kernels (the deterministic serial order of the queue atomics), n_threads for the intersections.
// pathCreateKernelWPT (cu:17-49) is restated here in its serial order (threadIdx / shared memory / atomicAdd make its text CUDA-only); one sample
// per pixel (the default uniform block sampler visits every block once per pass, BlockSamplerBuffer.h:38-46).  C++ leaves the evaluation order of
// the two randomFloat2() arguments of sampleSensorRay unspecified; device code evaluates left to right (jitter = 2-D dimension 0, aperture = 1),
// which is also PathTracer's explicit order (PathTracer.cu:186-188).  Returns the ray count (N per __internal__IntersectBuffers call).
unsigned long long ref_render_wavefront(const ctl_scene_view* view, int w, int h, int pass_first, int n_passes,
                                        int max_path_length, int rr_start, int direct, ctl_pixel_data* img, int n_threads, unsigned int* queue_sizes /* 2 * max_path_length or NULL: last pass */)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	pack_scene(*view);
	static bool sampler_init = false;
	if (!sampler_init) { new (&(*g_SamplerDataHost)) SamplerData(4096, 30); sampler_init = true; }
	delete g_gen; g_gen = new SamplingSequenceGeneratorHost<IndependantSamplingSequenceGenerator>();
	for (int p = 0; p < pass_first; p++) g_gen->Compute(*g_SamplerDataHost);
	g_wpt_threads = n_threads < 1 ? 1 : n_threads;
	Image I(w, h);
	I.Clear();
	for (int yy = 0; yy < h; yy++) for (int xx = 0; xx < w; xx++) memcpy((void*)&I.getPixelData(xx, yy), &img[yy * w + xx], sizeof(PixelData));
	WavefrontPathTracerBuffer* buf = new WavefrontPathTracerBuffer((unsigned)(w * h), (unsigned)(w * h));
	unsigned long long rays = 0;
	for (int p = 0; p < n_passes; p++) {
		g_gen->Compute(*g_SamplerDataHost);
		const int passes_done = pass_first + p + 1;      // m_uPassesDone++ precedes DoRender (Kernel/Tracer.h:231-232)
		buf->StartFrame(g_SceneData.m_rayTraceEps);
		memcpy((void*)&g_ray_buffer.As(), (void*)buf, sizeof(*buf));
		for (int rayidx = 0; rayidx < w * h; rayidx++) {
			int x = rayidx % w, y = rayidx / w;
			auto rng = g_SamplerData((unsigned)rayidx);
			NormalizedT<Ray> r;
			Vec2f jitter = rng.randomFloat2(), aperture = rng.randomFloat2();
			Spectrum W = g_SceneData.sampleSensorRay(r, Vec2f((float)x, (float)y) + jitter, aperture);
			WavefrontPTRayData dat;
			dat.x = half((float)x); dat.y = half((float)y); dat.throughput = W; dat.L = Spectrum(0.0f); dat.dIdx = UINT_MAX; dat.specular_bounce = true;
			dat.directF = Spectrum(0.0f); dat.dDist = 0; dat.bsdf_pdf = 0; dat.prev_normal = 0;   // left uninitialised by the reference; never read before written
			g_ray_buffer->insertPayloadElement(dat, r);
		}
		memcpy((void*)buf, (void*)&g_ray_buffer.As(), sizeof(*buf));
		int pass = 0;
		do {
			g_wpt_calls.clear();
			buf->FinishIteration();   // primary batch, then the secondary batch if there is one (DoubleRayBuffer.h:84-112)
			for (int n : g_wpt_calls) rays += (unsigned long long)n;   // g_RayTracedCounterHost += N per call (TraceHelper.cu:745)
			if (queue_sizes && p == n_passes - 1) { queue_sizes[2 * pass] = (unsigned)g_wpt_calls[0]; queue_sizes[2 * pass + 1] = g_wpt_calls.size() > 1 ? (unsigned)g_wpt_calls[1] : 0u; }
			memcpy((void*)&g_ray_buffer.As(), (void*)buf, sizeof(*buf));
			if (direct) pathIterateKernel<true>(I, pass, passes_done, max_path_length, rr_start, false);
			else pathIterateKernel<false>(I, pass, passes_done, max_path_length, rr_start, false);
			memcpy((void*)buf, (void*)&g_ray_buffer.As(), sizeof(*buf));
		} while (!buf->isEmpty() && ++pass < max_path_length);
	}
	buf->Free(); delete buf;
	for (int yy = 0; yy < h; yy++) for (int xx = 0; xx < w; xx++) memcpy(&img[yy * w + xx], (void*)&I.getPixelData(xx, yy), sizeof(PixelData));
	I.Free();
	return rays;
}

// traceRay on a batch (t in (rayEps, FLT_MAX)); out: ctl_trace_result
void ref_trace_rays(const ctl_scene_view* view, int n, const ctl_traversal_ray* rays, ctl_trace_result* out)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	pack_scene(*view);
	for (int i = 0; i < n; i++) {
		TraceResult r2 = traceRay(Ray(Vec3f(rays[i].o[0], rays[i].o[1], rays[i].o[2]), Vec3f(rays[i].d[0], rays[i].d[1], rays[i].d[2])));
		out[i].dist = r2.m_fDist; out[i].u = r2.m_fBaryCoords.x; out[i].v = r2.m_fBaryCoords.y; out[i].tri_idx = r2.m_triIdx; out[i].node_idx = r2.m_nodeIdx;
	}
}

// copySamplesToOutput (Kernel/ImagePipeline/ImagePipeline.cu:7-21) with the reference's PixelData::toSpectrum / toSRGB / toRGBCOL
void ref_resolve_srgb8(const ctl_pixel_data* img, int n, float splat_scale, unsigned char* rgba) {
	for (int i = 0; i < n; i++) {
		PixelData pd; memcpy((void*)&pd, &img[i], sizeof(pd));
		Spectrum c = pd.toSpectrum(splat_scale), c2;
		c.toSRGB(c2[0], c2[1], c2[2]);
		RGBCOL o = Spectrum(c2).toRGBCOL();
		rgba[4 * i] = o.x; rgba[4 * i + 1] = o.y; rgba[4 * i + 2] = o.z; rgba[4 * i + 3] = o.w;
	}
}

// applyImagePipeline(tracer, img, filter): rtm_Copy (CanonicalFilter.cu:29-36) + copyFilteredToOutput (ImagePipeline.cu:32-41)
// with the reference's evalFilter, Filter aggregate, toRGBE / fromRGBE / toSRGB / toRGBCOL
void ref_resolve_filtered_srgb8(const ctl_pixel_data* img, int w, int h, float splat_scale, int type, float xw, float yw, float alpha, unsigned char* rgba) {
	std::vector<PixelData> P((size_t)w * h);
	memcpy((void*)P.data(), img, (size_t)w * h * sizeof(PixelData));
	Filter filter;
	if (type == 0) filter.SetData(BoxFilter(xw, yw)); else if (type == 1) filter.SetData(GaussianFilter(xw, yw, alpha)); else filter.SetData(TriangleFilter(xw, yw));
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
		Spectrum c = evalFilter(filter, P.data(), splat_scale, x, y, w, h);
		RGBE e = c.toRGBE();
		Spectrum s; s.fromRGBE(e);
		Spectrum c2; s.toSRGB(c2[0], c2[1], c2[2]);
		RGBCOL o = Spectrum(c2).toRGBCOL();
		int i = y * w + x;
		rgba[4 * i] = o.x; rgba[4 * i + 1] = o.y; rgba[4 * i + 2] = o.z; rgba[4 * i + 3] = o.w;
	}
}

// applyImagePipeline(tracer, img, filter, process) in full (ImagePipeline.cu:54-84) with the reference's own per-pixel code: evalFilter over the
// Filter aggregate (all five filters), toRGBE / fromRGBE, getLuminance, toYxy / fromYxy, toRGBCOL / fromRGBCOL, toSRGB.  The two kernels that
// are CUDA-only text (computeLuminanceInfo's atomics, Reinhard05Kernel's thread indexing) are looped here: block by block, row-major inside a
// 16x16 block for the luminance sums (the reference's order is the scheduler's).
void ref_apply_image_pipeline(const ctl_pixel_data* img, int w, int h, float splat_scale, const ctl_image_pipeline* P, unsigned char* rgba, float* lum_out) {
	std::vector<PixelData> px((size_t)w * h);
	memcpy((void*)px.data(), img, (size_t)w * h * sizeof(PixelData));
	auto gamma_out = [&](const Spectrum& c, int i) {
		Spectrum c2; c.toSRGB(c2[0], c2[1], c2[2]);
		RGBCOL o = Spectrum(c2).toRGBCOL();
		rgba[4 * i] = o.x; rgba[4 * i + 1] = o.y; rgba[4 * i + 2] = o.z; rgba[4 * i + 3] = o.w;
	};
	if (P->filter_type < 0 && !P->tonemap) { for (int i = 0; i < w * h; i++) gamma_out(px[i].toSpectrum(splat_scale), i); return; }
	std::vector<RGBE> stage2((size_t)w * h);
	if (P->filter_type >= 0) {
		Filter filter;
		switch (P->filter_type) {
		case 0: filter.SetData(BoxFilter(P->x_width, P->y_width)); break;
		case 1: filter.SetData(GaussianFilter(P->x_width, P->y_width, P->param0)); break;
		case 2: filter.SetData(TriangleFilter(P->x_width, P->y_width)); break;
		case 3: filter.SetData(MitchellFilter(P->param0, P->param1, P->x_width, P->y_width)); break;
		default: filter.SetData(LanczosSincFilter(P->x_width, P->y_width, P->param0)); break;
		}
		for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) stage2[(size_t)y * w + x] = evalFilter(filter, px.data(), splat_scale, x, y, w, h).toRGBE();
	} else for (int i = 0; i < w * h; i++) stage2[i] = px[i].toSpectrum(splat_scale).toRGBE();
	if (!P->tonemap) { for (int i = 0; i < w * h; i++) { Spectrum s; s.fromRGBE(stage2[i]); gamma_out(s, i); } return; }
	float mn = FLT_MAX, mx = 0.0f, sum = 0.0f, sumlog = 0.0f;
	for (int by = 0; by < h; by += 16) for (int bx = 0; bx < w; bx += 16) {
		float sb = 0.0f, sl = 0.0f;
		for (int y = by; y < std::min(h, by + 16); y++) for (int x = bx; x < std::min(w, bx + 16); x++) {
			Spectrum L_w; L_w.fromRGBE(stage2[(size_t)y * w + x]);
			float Y = L_w.getLuminance();
			mn = std::min(mn, Y); mx = std::max(mx, Y); sb += Y; sl += math::log(2.3e-5f + Y);
		}
		sum += sb; sumlog += sl;
	}
	float avgLum = sum / (w * h), logAvgLuminance = math::exp(sumlog / (w * h));
	ToneMapPostProcess tm; tm.m_key = P->key; tm.m_burn = P->burn;
	float scale = tm.m_key / logAvgLuminance, Lwhite = mx * scale;               // ToneMapPostProcess.cu:33-36
	auto burn = min(1.0f, max(1e-8f, 1.0f - tm.m_burn));
	float invWp2 = 1 / (Lwhite * Lwhite * std::pow(burn, 4.0f));
	if (lum_out) { lum_out[0] = mn; lum_out[1] = mx; lum_out[2] = avgLum; lum_out[3] = logAvgLuminance; lum_out[4] = scale; lum_out[5] = invWp2; }
	for (int i = 0; i < w * h; i++) {
		Spectrum color; color.fromRGBE(stage2[i]);                                 // Reinhard05Kernel body, ToneMapPostProcess.cu:11-22
		float x, y, Y; color.toYxy(Y, x, y);
		float Lp = scale * Y;
		Y = Lp * (1.0f + Lp * invWp2) / (1.0f + Lp);
		color.fromYxy(Y, x, y);
		RGBCOL processed = color.toRGBCOL();
		Spectrum s; s.fromRGBCOL(processed);                                       // applyGammaCorrectureToOutput, ImagePipeline.cu:43-52
		gamma_out(s, i);
	}
}

// PixelVarianceBuffer::AddPass with the uniform block sampler (every block flag = 1): the reference's own PixelVarianceInfo::updateMoments
void ref_variance_add_pass(ctl_pixel_variance_info* var, const ctl_pixel_data* img, int n, float splat_scale) {
	static_assert(sizeof(PixelVarianceInfo) == sizeof(ctl_pixel_variance_info), "PixelVarianceInfo layout");
	for (int i = 0; i < n; i++) {
		PixelVarianceInfo V; memcpy((void*)&V, &var[i], sizeof(V));
		PixelData pd; memcpy((void*)&pd, &img[i], sizeof(pd));
		V.updateMoments(pd, splat_scale, 1.0f);
		memcpy(&var[i], (void*)&V, sizeof(V));
	}
}

// NonLocalMeansFilter::Apply (Kernel/ImagePipeline/Filter/NonLocalMeansFilter.cu:184-228) with the reference's OWN kernels (lines 9-159 of that file, compiled
// as host functions, build_ref.sh note 13).  The launch geometry is the reference's: copyToCached and applyWeights over (w/16+1, h/16+1) blocks of 16x16,
// computeWeights in 200-pixel super-blocks of 13x13 blocks with (x_off, y_off).  A block's threads run twice: the first sweep completes the block's 64x64
// tile cache (copyToShared is per thread), the second computes from the complete cache -- what __syncthreads guarantees on the device; outputs are overwrites.
// weights: w*h*169 floats owned by the caller (the filter's m_weightBuffer); recomputed when compute_weights != 0, else applied as they are.
void ref_nlm_filter(const ctl_pixel_data* img, const ctl_pixel_variance_info* var, int w, int h, float splat_scale, float k, float sigma2Scale,
                               float* weights, int compute_weights, unsigned char* rgbe_out)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	const int R = 6, F = 3;
	Image I(w, h);
	PixelVarianceBuffer vb(w, h);
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
		memcpy((void*)&I.getPixelData(x, y), &img[y * w + x], sizeof(PixelData));
		memcpy((void*)&vb(x, y), &var[y * w + x], sizeof(PixelVarianceInfo));
	}
	std::vector<RGBE> cached((size_t)w * h);
	NonLocalMeansFilter::FilterWeightBuffer wb;
	wb.R = R; wb.F = F; wb.w = w; wb.h = h; wb.n_weights_per_pixel = (2 * R + 1) * (2 * R + 1); wb.deviceWeightBuffer = weights;
	nlm_launch(w / 16 + 1, h / 16 + 1, [&]() { copyToCached(I, cached.data(), splat_scale); });
	if (compute_weights) {
		memset(weights, 0, (size_t)w * h * wb.n_weights_per_pixel * sizeof(float));   // m_weightBuffer.ClearBuffer()
		const int block_width = 200;
		const int n_blocks_x = (w + block_width - 1) / block_width, n_blocks_y = (h + block_width - 1) / block_width, n_cuda_blocks = (block_width + 16 - 1) / 16;
		for (int i = 0; i < n_blocks_x; i++) for (int j = 0; j < n_blocks_y; j++)
			nlm_launch(n_cuda_blocks, n_cuda_blocks, [&]() { computeWeights(I, cached.data(), 0, R, F, k, sigma2Scale, vb, wb, i * block_width, j * block_width); });
	}
	nlm_launch(w / 16 + 1, h / 16 + 1, [&]() { applyWeights(I, cached.data(), wb, R, F); });
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) memcpy(rgbe_out + 4 * ((size_t)y * w + x), (void*)&I.getFilteredData(x, y), 4);
	wb.deviceWeightBuffer = 0;
	vb.Free(); I.Free();
}

// .xmsh WRITER: the reference's own Mesh::CompileMesh (Engine/Mesh.cpp:199-290: TriangleData encoding, vertex normals, MeshPartLight list,
// Material blobs, SplitBVHBuilder via ConstructBVH) behind the MeshCompileType token the scene loader expects first (Engine/DynamicScene.cpp:313-319,
// Engine/MeshLoader/MeshCompiler.cpp:94).  sub_tris[k] = triangles of sub-mesh k (consecutive); materials / emissive per sub-mesh.
int ref_write_xmsh(const char* path, const float* verts, unsigned nv, const unsigned* indices, unsigned n_indices, const unsigned* sub_tris, unsigned n_sub,
                   const ctl_material* mats, const float* emissive)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	try {
		ctl_scene_view v; memset(&v, 0, sizeof(v)); v.materials = mats; v.n_materials = n_sub; v.camera.resolution[0] = v.camera.resolution[1] = 16;
		v.camera.inv_resolution[0] = v.camera.inv_resolution[1] = 1.0f / 16; for (int i = 0; i < 4; i++) { v.camera.sample_to_camera[i * 5] = 1; v.camera.to_world[i * 5] = 1; }
		pack_scene(v);   // builds the reference Material objects (g_scene->mats) from the compact records
		std::vector<Material> M = g_scene->mats; std::vector<Spectrum> Le(n_sub);
		for (unsigned k = 0; k < n_sub; k++) {
			M[k].Name = FixedString<64>(format("material_%u", k)); M[k].NodeLightIndex = UINT_MAX;
			Le[k] = emissive ? Spectrum(emissive[3 * k], emissive[3 * k + 1], emissive[3 * k + 2]) : Spectrum(0.0f);
		}
		FileOutputStream out(path);
		out << (unsigned int)MeshCompileType::Static;
		Mesh::CompileMesh((const Vec3f*)verts, nv, 0, 0, 0, indices, n_indices, M.data(), Le.data(), sub_tris, 0, out, false, false, 0.0f);
		out.Close();
	} catch (const std::exception& e) { fprintf(stderr, "ref_write_xmsh: %s\n", e.what()); return 1; }
	return 0;
}

// OBJ / PLY -> .xmsh with the reference's OWN mesh compilers (Engine/MeshLoader/ObjParser.cpp compileobj, PlyParser.cpp compileply), framed like
// MeshCompilerManager::Compile (Engine/MeshLoader/MeshCompiler.cpp:82-99: the MeshCompileType token first)
int ref_compile_mesh(const char* in_path, const char* xmsh_path)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	try {
		std::string p(in_path);
		const bool obj = p.size() > 4 && p.substr(p.size() - 4) == ".obj", ply = p.size() > 4 && p.substr(p.size() - 4) == ".ply";
		if (!obj && !ply) throw std::runtime_error("ref_compile_mesh: .obj or .ply expected");
		IInStream* in = OpenFile(p);
		FileOutputStream out(xmsh_path);
		out << (unsigned int)MeshCompileType::Static;
		if (obj) compileobj(*in, out); else compileply(*in, out);
		out.Close();
		delete in;
	} catch (const std::exception& e) { fprintf(stderr, "ref_compile_mesh: %s\n", e.what()); return 1; }
	return 0;
}

// layout facts the product's .xmsh reader hard-codes (cudatracerlib_b200/csrc/xmsh.cpp), checked against the reference headers
static_assert(sizeof(Material) == 3344 && offsetof(Material, NodeLightIndex) == 68 && offsetof(Material, bsdf) == 512 && sizeof(FixedString<64>) == 68, "Material layout");
static_assert(offsetof(Material, usedBssrdf) == 88 && offsetof(Material, NormalMap) == 2656 && offsetof(Material, HeightMap) == 2880 && offsetof(Material, AlphaMap) == 3104 &&
              offsetof(AlphaBlendData, state) == 0 && offsetof(Material::mpHlp, used) == 0, "Material map-usage fields");
static_assert(sizeof(MeshPartLight) == 48 && offsetof(MeshPartLight, L) == 36 && sizeof(Texture) == 208 && sizeof(AABB) == 24 && sizeof(BSDF) == 56, "xmsh record layout");
static_assert(offsetof(BSDF, m_enableTwoSided) == 52 && offsetof(diffuse, m_reflectance) == 64 && offsetof(ConstantTexture, val) == 8, "BSDF layout");
static_assert(offsetof(roughconductor, m_specularReflectance) == 64 && offsetof(roughconductor, m_alphaU) == 272 && offsetof(roughconductor, m_alphaV) == 480 &&
              offsetof(roughconductor, m_eta) == 688 && offsetof(roughconductor, m_k) == 700 && offsetof(roughconductor, m_type) == 716, "roughconductor layout");
static_assert(offsetof(dielectric, eta_f) == 64 && offsetof(dielectric, m_specularTransmittance) == 128 && offsetof(dielectric, m_specularReflectance) == 336 &&
              offsetof(DispersionCauchy, B) == 8 && offsetof(DispersionCauchy, C) == 12 && sizeof(Dispersion) == 64, "dielectric layout");
static_assert(diffuse::TYPE() == 1 && dielectric::TYPE() == 3 && roughconductor::TYPE() == 7 && ConstantTexture::TYPE() == 2 && DispersionCauchy::TYPE() == 1 &&
              (unsigned)MeshCompileType::Static == 0, "type tags");

// Known-answer probes of the reference's own math (regenerates SURVEY Appendix C)
void ref_xorwow_floats(unsigned int seed_subsequence, int n, float* out) { CudaRNG rng(seed_subsequence); for (int i = 0; i < n; i++) out[i] = rng.randomFloat(); }
void ref_woop_setdata(const float* v0, const float* v1, const float* v2, float* out12) {
	TriIntersectorData T; T.setData(Vec3f(v0[0], v0[1], v0[2]), Vec3f(v1[0], v1[1], v1[2]), Vec3f(v2[0], v2[1], v2[2])); memcpy(out12, &T, 48);
}
void ref_sample_tables(unsigned int pass, float* d1, float* d2) {
	std::lock_guard<std::mutex> lock(g_mutex);
	static bool sampler_init = false;
	SamplingSequenceGeneratorHost<IndependantSamplingSequenceGenerator> gen;
	static SamplerData* data = nullptr;
	if (!data) data = new SamplerData(4096, 30);
	for (unsigned p = 0; p <= pass; p++) gen.Compute(*data);
	for (unsigned s = 0; s < 4096; s++) for (unsigned i = 0; i < 30; i++) {
		d1[i * 4096 + s] = data->getSequenceElement1(s, i);
		Vec2f q = data->getSequenceElement2(s, i); d2[(i * 4096 + s) * 2] = q.x; d2[(i * 4096 + s) * 2 + 1] = q.y;
	}
	(void)sampler_init;
}
// BSDFALL::sample / f / pdf of a ctl_material at a local wi (identity frame): out9 = weight rgb, pdf, wo xyz, sampledType, eta
void ref_bsdf_probe(const ctl_material* m, const float* wi, float sx, float sy, float* out9, float* f3, float* pdf1) {
	std::lock_guard<std::mutex> lock(g_mutex);
	ctl_scene_view v; memset(&v, 0, sizeof(v)); v.materials = m; v.n_materials = 1; v.camera.resolution[0] = v.camera.resolution[1] = 16;
	v.camera.inv_resolution[0] = v.camera.inv_resolution[1] = 1.0f / 16; for (int i = 0; i < 4; i++) { v.camera.sample_to_camera[i * 5] = 1; v.camera.to_world[i * 5] = 1; }
	pack_scene(v);
	BSDFSamplingRecord bRec;
	DifferentialGeometry& dg = bRec.dg;
	dg.P = Vec3f(0); dg.sys = Frame(NormalizedT<Vec3f>(1, 0, 0), NormalizedT<Vec3f>(0, 1, 0), NormalizedT<Vec3f>(0, 0, 1)); dg.n = NormalizedT<Vec3f>(0, 0, 1); dg.uv[0] = Vec2f(0);
	bRec.wi = NormalizedT<Vec3f>(wi[0], wi[1], wi[2]); bRec.mode = ERadiance; bRec.typeMask = EAll; bRec.sampledType = 0; bRec.eta = 1.0f;
	float pdf = 0;
	Spectrum w = g_scene->mats[0].bsdf.sample(bRec, pdf, Vec2f(sx, sy));
	float r, g, b; w.toLinearRGB(r, g, b);
	out9[0] = r; out9[1] = g; out9[2] = b; out9[3] = pdf; out9[4] = bRec.wo.x; out9[5] = bRec.wo.y; out9[6] = bRec.wo.z; out9[7] = (float)bRec.sampledType; out9[8] = bRec.eta;
	bRec.typeMask = EAll;
	Spectrum f = g_scene->mats[0].bsdf.f(bRec); f.toLinearRGB(f3[0], f3[1], f3[2]);
	*pdf1 = g_scene->mats[0].bsdf.pdf(bRec);
}

} // extern "C"
