/*
 * ctl_b200.h -- C ABI of the B200-native replacement for CudaTracerLib's
 * ray-traversal + path-tracing hot path.
 *
 * Every entry point cites the reference interface (file:line, relative to the
 * CudaTracerLib tree) that it replaces.  Plain pointers and sizes only; no C++
 * types, no torch types.  All functions returning int use 0 = ok, non-zero =
 * error; the message is available from ctl_last_error() (the C++ adapter in
 * cudatracerlib_b200/csrc/b200_path_tracer.h re-raises it as
 * std::runtime_error, the reference's ThrowCudaErrors convention,
 * Defines.cpp:15-29).
 *
 * The data surface (ctl_scene_view) mirrors the hot subset of
 * KernelDynamicScene (Engine/KernelDynamicScene.h:28-57) byte for byte where
 * the reference layout is on the path (BVH nodes, Woop triangles, leaf index
 * words, TriangleData, KernelMesh, Node, float4x4, PixelData, traversalRay,
 * traversalResult) and replaces the 3344-byte tagged-union Material / 592-byte
 * Light by compact 64-byte records that carry exactly the fields the path
 * reads.
 */
#ifndef CTL_B200_H
#define CTL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTL_MAX_NUM_LIGHTS 16 /* Engine/KernelDynamicScene.h:26 */
#define CTL_SENTINEL 0x76543210 /* Kernel/TraceHelper.cu:20 */

/* ---- byte-exact reference layouts ------------------ */

/* Engine/TriIntersectorData.h:42-117; 64 B; children in float4 units */
typedef struct ctl_bvh_node {
    float a[4]; /* c0.lo.x c0.hi.x c0.lo.y c0.hi.y */
    float b[4]; /* c1.lo.x c1.hi.x c1.lo.y c1.hi.y */
    float c[4]; /* c0.lo.z c0.hi.z c1.lo.z c1.hi.z */
    int32_t child0, child1;
    uint32_t parent;
    uint32_t pad;
} ctl_bvh_node;

/* Engine/TriIntersectorData.h:30-40; 48 B */
typedef struct ctl_woop_tri {
    float a[4], b[4], c[4];
} ctl_woop_tri;

/* Engine/TriangleData.h:19-35 (EXT_TRI, NUM_UV_SETS 1); 32 B */
typedef struct ctl_tri_data {
    uint32_t w[8];
} ctl_tri_data;

/* Engine/Mesh.h:12-19; 20 B */
typedef struct ctl_mesh {
    uint32_t tri_offset;      /* m_uTriangleOffset                     */
    uint32_t bvh_node_offset; /* m_uBVHNodeOffset, float4 units        */
    uint32_t bvh_tri_offset;  /* m_uBVHTriangleOffset, float4 units    */
    uint32_t bvh_idx_offset;  /* m_uBVHIndicesOffset, slots            */
    uint32_t mat_offset;      /* m_uStdMaterialOffset                  */
} ctl_mesh;

/* SceneTypes/Node.h:13-20 with FixedSizeArray<unsigned,2,true,0xff> m_uLights = {length, buffer[2]}
 * (Base/FixedSizeArray.h:107-109); 24 B.  Layout checked against the reference headers by oracle/ref_driver.cpp. */
typedef struct ctl_node {
    uint32_t mesh_index;
    uint32_t material_offset;
    uint32_t instanciated_material;
    uint32_t n_lights;
    uint32_t lights[2]; /* 0xffffffff = none */
} ctl_node;

/* Kernel/TraceHelper.h:55-59; 32 B */
typedef struct ctl_traversal_ray {
    float o[3], tmin;
    float d[3], tmax;
} ctl_traversal_ray;

/* Kernel/TraceHelper.h:61-69; 16 B. miss = (dist bits, -1, -1, 0) */
typedef struct ctl_traversal_result {
    float dist;
    int32_t node_idx;
    int32_t tri_idx;
    uint32_t bary; /* (u16)(v*65535) << 16 | (u16)(u*65535) */
} ctl_traversal_result;

/* Kernel/TraceResult.h:17-31 (field order dist,u,v,tri,node); 24 B padded to 32 B
 * is NOT used on the wire; this 24 B form is what ctl_trace_rays writes. */
typedef struct ctl_trace_result {
    float dist;
    float u, v;
    uint32_t tri_idx;  /* UINT_MAX = miss */
    uint32_t node_idx;
} ctl_trace_result;

/* Engine/Image.h:10-29; 28 B */
typedef struct ctl_pixel_data {
    float rgb[3];
    float rgb_splat[3];
    float weight_sum;
} ctl_pixel_data;

/* Engine/ShapeSet.h:16-26; 64 B */
typedef struct ctl_light_tri {
    float p[3][3];
    float n[3];
    float area;
    uint32_t i_dat;
    uint32_t t_dat;
    uint32_t pad;
} ctl_light_tri;

/* Kernel/PixelVarianceBuffer.h:10-67 (PixelVarianceInfo: prev_I, half_buffer, iterations_done, weight, VarAccumulator<float> I,
 * num_samples_var); 44 B */
typedef struct ctl_pixel_variance_info {
    float prev_I[3];
    float half_buffer[3];
    int32_t iterations_done;
    float weight;
    float sum_x, sum_x2;
    int32_t num_samples_var;
} ctl_pixel_variance_info;

/* applyImagePipeline's two optional stages (Kernel/ImagePipeline/ImagePipeline.h): an ImageSamplesFilter (CanonicalFilter over one of
 * the five filters of SceneTypes/Filter.h) and a PostProcess (ToneMapPostProcess = Reinhard05). */
typedef struct ctl_image_pipeline {
    int32_t filter_type; /* -1 none; 0 BoxFilter, 1 GaussianFilter, 2 TriangleFilter, 3 MitchellFilter, 4 LanczosSincFilter (CanonicalFilter);
                          *  5 NonLocalMeansFilter (Kernel/ImagePipeline/Filter/NonLocalMeansFilter.h) */
    float x_width, y_width; /* NonLocalMeansFilter: x_width = UpdateWeightPeriodicity (integer >= 1, reference default 25), y_width unused */
    float param0, param1; /* Gaussian: alpha; Mitchell: B, C; LanczosSinc: tau; NonLocalMeans: k (0.45), sigma2Scale (0.005) */
    int32_t tonemap;     /* 0 none; 1 ToneMapPostProcess (Kernel/ImagePipeline/PostProcess/ToneMapPostProcess.h) */
    float key, burn;     /* m_key (0.18), m_burn (0) */
} ctl_image_pipeline;

/* ---- compact records replacing Material / Light ------------------------ */

enum { CTL_BSDF_DIFFUSE = 0, CTL_BSDF_ROUGHCONDUCTOR = 1, CTL_BSDF_DIELECTRIC = 2 };
enum { CTL_DISTR_BECKMANN = 0, CTL_DISTR_GGX = 1 };
#define CTL_MAT_TWO_SIDED 1u

/* Fields read by the path from Material (Engine/Material.h:38-112): NodeLightIndex,
 * bsdf type + payload (SceneTypes/BSDF_Simple.h), m_enableTwoSided. 64 B. */
typedef struct ctl_material {
    uint32_t bsdf_type;        /* CTL_BSDF_*                                      */
    uint32_t flags;            /* CTL_MAT_TWO_SIDED                               */
    uint32_t node_light_index; /* Material::NodeLightIndex, 0xffffffff = none     */
    uint32_t distr_type;       /* CTL_DISTR_* (roughconductor)                    */
    float reflectance[3];      /* diffuse m_reflectance / specularReflectance     */
    float alpha_u;
    float eta[3];              /* conductor eta; eta[0] = dielectric eta(600 nm)  */
    float alpha_v;
    float k[3];                /* conductor k                                     */
    float transmittance;       /* dielectric specular transmittance (grey)        */
} ctl_material;

/* DiffuseLight (SceneTypes/Light.h:96-142) over a ShapeSet (Engine/ShapeSet.h). 32 B */
typedef struct ctl_light {
    float radiance[3];
    float sum_area;
    uint32_t tri_offset; /* first ctl_light_tri                              */
    uint32_t cdf_offset; /* first of count+1 floats in light_cdf_data        */
    uint32_t count;
    uint32_t node_idx;
} ctl_light;

/* PerspectiveSensor (SceneTypes/Sensor.h:189, Sensor.cu:76-144) */
typedef struct ctl_camera {
    float sample_to_camera[16]; /* row-major float4x4 */
    float to_world[16];
    float inv_resolution[2];
    float resolution[2];
} ctl_camera;

/* Flat, read-only view of everything the path reads. Host pointers; copied
 * by ctl_upload_scene.  == KernelDynamicScene's hot subset
 * (Engine/KernelDynamicScene.h:28-57, Engine/DynamicScene.cpp:567-589). */
typedef struct ctl_scene_view {
    const ctl_bvh_node* bvh_nodes;       uint32_t n_bvh_nodes;       /* m_sBVHNodeData  */
    const ctl_woop_tri* woop;            uint32_t n_woop;            /* m_sBVHIntData   */
    const uint32_t*     tri_index;       uint32_t n_tri_index;       /* m_sBVHIndexData */
    const ctl_tri_data* tri_data;        uint32_t n_tri_data;        /* m_sTriData      */
    const ctl_mesh*     meshes;          uint32_t n_meshes;          /* m_sMeshData     */
    const ctl_node*     nodes;           uint32_t n_nodes;           /* m_sNodeData     */
    const float*        node_xf;         /* n_nodes * 16, m_pNodeTransforms           */
    const float*        node_inv_xf;     /* n_nodes * 16, m_pInvNodeTransforms        */
    const ctl_bvh_node* scene_bvh_nodes; uint32_t n_scene_bvh_nodes; /* m_sSceneBVH    */
    int32_t             scene_start_node;                            /* m_sStartNode   */
    const ctl_material* materials;       uint32_t n_materials;       /* m_sMatData     */
    const ctl_light*    lights;          uint32_t n_lights_buf;      /* m_sLightBuf    */
    const ctl_light_tri* light_tris;     uint32_t n_light_tris;      /* m_sAnimData    */
    const float*        light_cdf_data;  uint32_t n_light_cdf_data;  /* m_sAnimData    */
    uint32_t num_lights;                                             /* m_numLights    */
    uint32_t light_indices[CTL_MAX_NUM_LIGHTS];                      /* m_pLightIndices*/
    float    light_cdf[CTL_MAX_NUM_LIGHTS];                          /* m_pLightCDF    */
    ctl_camera camera;                                               /* m_Camera       */
    float box_min[3], box_max[3];                                    /* m_sBox         */
    float ray_eps;                                                   /* m_rayTraceEps  */
    const uint32_t* node_alias; /* NULL, or n_nodes entries when the scene level was re-braided (ctl_scene_set_rebraid): the instance each (pseudo-)node stands for */
} ctl_scene_view;

/* ---- host-side scene construction (replaces DynamicScene + SplitBVHBuilder
 *      for synthetic scenes; Engine/DynamicScene.cpp, BVHBuilderHelper.cpp) --- */

typedef struct ctl_scene ctl_scene;

/* kind: 0 = Cornell-32 single node, 1 = Cornell-32 split in 7 nodes,
 *       2 = C2 100K diffuse, 3 = C3 100K microfacet mix,
 *       4 = C4 1M clustered + foliage (8 meshes/nodes), 5 = C5 (C4 geometry + C3 materials),
 *       6 = small random "soup" test scene (n_hint triangles).                        */
ctl_scene* ctl_scene_create(int kind, int width, int height, uint32_t seed, int n_hint);
/* Build from caller geometry: one mesh, one node, identity transform.
 * verts: nv*3 floats; indices: nt*3; mat_index: nt bytes; materials: nm records;
 * emissive[nm*3]: radiance per material (all-zero = not a light). */
ctl_scene* ctl_scene_create_from_mesh(const float* verts, uint32_t nv, const uint32_t* indices, uint32_t nt,
                                      const uint8_t* mat_index, const ctl_material* materials, uint32_t nm,
                                      const float* emissive, const float* cam_pos, const float* cam_target,
                                      const float* cam_up, float fov_deg, int width, int height);
/* Flat scene import from the reference's compiled-mesh files (.xmsh; format in csrc/xmsh.cpp): one mesh + one node per file, == the static-mesh
 * branch of DynamicScene::CreateNode (Engine/DynamicScene.cpp:283-345) over Mesh::Mesh(IInStream&) (Engine/Mesh.cpp:46-98).  The file's BVH
 * nodes / Woop triangles / leaf words / TriangleData are used as they are (reference layouts); its Material blobs are decoded to ctl_material
 * (diffuse, roughconductor, dielectric with constant textures; anything else fails with a message naming the material); MeshPartLight entries
 * become area lights.  node_xforms: n_files row-major float4x4 local-to-world matrices, or NULL = identity. */
ctl_scene* ctl_scene_create_from_xmsh(const char* const* paths, uint32_t n_files, const float* node_xforms, const float* cam_pos,
                                      const float* cam_target, const float* cam_up, float fov_deg, int width, int height);
/* Same import for any mix of mesh files, chosen by extension like the reference's MeshCompilerManager (Engine/MeshLoader/MeshCompiler.cpp):
 * .xmsh as above; .obj (+ the .mtl it names) through this library's own OBJ front end, which reproduces what the reference's
 * compileobj -> Mesh::CompileMesh writes for the same file (vertex de-duplication order, fan triangulation, reversed winding, its
 * single-precision number reader, vertex normals, UV-driven dpdu / dpdv) for materials of the hot path: illum 2 with Ks = 0 (diffuse),
 * illum 7 / 9 (dielectric), Ke (area light); .ply (ascii, binary little / big endian; float x y z vertices, triangle / quad faces) likewise
 * reproduces compileply (Engine/MeshLoader/PlyParser.cpp:182-371).  ctl_scene_create_from_xmsh accepts the same mix (it is this function). */
ctl_scene* ctl_scene_create_from_files(const char* const* paths, uint32_t n_files, const float* node_xforms, const float* cam_pos,
                                       const float* cam_target, const float* cam_up, float fov_deg, int width, int height);
/* Mesh `mesh` of a host scene as an .xmsh file: the output sequence of Mesh::CompileMesh (Engine/Mesh.cpp:278-289). */
int  ctl_scene_write_xmsh(const ctl_scene*, uint32_t mesh, const char* path);
/* Source triangles of mesh `mesh` of a host scene built here (9 floats per triangle, in TriangleData order); verts9_out may be NULL to query
 * *n_tris.  For export / BVH-rebuild tooling (the reference keeps them only inside its mesh compilers, Engine/Mesh.cpp:199-290). */
int  ctl_scene_get_mesh_triangles(const ctl_scene*, uint32_t mesh, float* verts9_out, uint32_t* n_tris);
/* == DynamicScene::SetNodeTransform (Engine/DynamicScene.cpp:433-443) on a host scene: new row-major local-to-world float4x4 of instance `node`.  The
 * node level is re-assembled (scene-level BVH as BVHRebuilder would, inverse matrices, the node's area lights as RecomputeShape would, scene box,
 * ray epsilon); mesh BVHs / Woop triangles / TriangleData are untouched.  Views obtained before the call are invalidated. */
int  ctl_scene_set_node_transform(ctl_scene*, uint32_t node, const float* xf16);
/* == DynamicScene::getKernelSceneData(false) (Engine/DynamicScene.cpp:567-589): the flat view of the host arrays; valid until the scene is changed or destroyed */
/* Partial re-braiding of the scene level (no reference counterpart; its BVHRebuilder keeps one leaf per instance): instances with large,
 * overlapping boxes are opened into up to max_entries (instance, sub-tree) leaves, each an ordinary node + mesh record over a re-based copy of the
 * sub-tree, so the traversal kernels and the data layout are unchanged.  Same hits; ctl_intersect / ctl_intersect_host / ctl_trace_rays_host report the
 * INSTANCE a hit belongs to (results pass through view.node_alias on the device), like the reference would.  0 = off; the default is by scene size (1 024 leaves when the meshes hold >= 32 768 BVH nodes, else off; CTL_REBRAID=<n> in the environment overrides).  Re-assembles the node level: obtain the view again, then ctl_upload_scene
 * (ctl_update_scene_nodes uploads everything for a re-braided view: its mesh-level records follow the node level). */
int  ctl_scene_set_rebraid(ctl_scene*, uint32_t max_entries);
int  ctl_scene_get_view(const ctl_scene*, ctl_scene_view* out);
/* == DynamicScene::~DynamicScene (Engine/DynamicScene.cpp:219) */
void ctl_scene_destroy(ctl_scene*);
/* Structural check of a view that was not built by this library (hand-filled arrays, data read from elsewhere) before ctl_upload_scene: child / leaf
 * references inside their arrays, each BVH a tree, every leaf run terminated, triangle / mesh / material / light indices in range, tree depths within the
 * 64-entry traversal stack.  The reference trusts its builders and would read out of bounds or spin in __traceRay_internal__ (Kernel/TraceHelper.cu:88-172)
 * on such data; the .xmsh reader runs the mesh part of this check on every file.  0 = consistent; otherwise ctl_last_error names the first problem. */
int  ctl_validate_scene_view(const ctl_scene_view*);
/* GPU construction of one mesh BVH in the reference layout -- replaces the CPU pre-process SplitBVHBuilder.cpp:163-203 /
 * BVHBuilderHelper.cpp:119 for meshes that need (re)building at run time.  Triangles are sorted by Morton code
 * (hand-written radix sort); the tree over them is built by parallel locally-ordered clustering (every merge decided by surface
 * area, the Morton order only bounds the search: csrc/bvh_ploc.cuh), then a bottom-up fit with the SAH leaf / split decision
 * (<= 8-triangle leaves), leaves laid out in tree order, nodes in pre-order.  CTL_GPU_BUILDER=lbvh selects the Karras radix tree
 * instead (fastest build, 1.1-1.3x slower traversal), CTL_PLOC_RADIUS the search window (default 16).  verts9: n_tris * 9 floats
 * (host).  Outputs (host): nodes_out (capacity max(1, n_tris)), woop_out / index_out (n_tris each, leaf order), *n_nodes_out,
 * optional device build time.  Deterministic. */
int ctl_bvh_build_gpu(int device, const float* verts9, uint32_t n_tris, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out,
                      ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms);
/* The same with the builder named: algorithm 0 = LBVH, 1 = agglomerative (PLOC), 2 = both (the lower SAH cost wins); radius <= 0 = default (16, at most 32). */
int ctl_bvh_build_gpu_ex(int device, const float* verts9, uint32_t n_tris, int algorithm, int radius, ctl_bvh_node* nodes_out,
                         uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms);
/* ... and with triangle pre-splitting (csrc/bvh_presplit.cuh): triangles whose box is much larger than the triangle -- long thin diagonals, where the
 * reference's SplitBVHBuilder (SplitBVHBuilder.cpp:205-330) takes spatial splits -- are replaced by up to 16 references with the tight boxes of their
 * clipped pieces before the tree is built; at most (1 + max_growth) * n_tris references in all (max_growth in [0, 8]; 0 = none).  The output arrays hold
 * `capacity` (>= n_tris) entries each; *n_slots_out = references written to woop_out / index_out (a triangle may appear in several leaves). */
int ctl_bvh_build_gpu_split(int device, const float* verts9, uint32_t n_tris, int algorithm, int radius, float max_growth, uint32_t capacity,
                            ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, uint32_t* n_slots_out, float* build_ms);
/* Rebuild all mesh BVHs of a host scene on the GPU (ctl_bvh_build_gpu_split; CTL_GPU_BUILDER, CTL_PLOC_RADIUS, CTL_GPU_SPLIT = reference growth budget,
 * default 3, 0 = no pre-splitting).  Views obtained before the call are invalidated. */
int ctl_scene_rebuild_bvh_gpu(ctl_scene*, int device, float* build_ms_total);
/* encoders exposed for known-answer tests */
void ctl_encode_woop(const float v0[3], const float v1[3], const float v2[3], ctl_woop_tri* out); /* TriIntersectorData.cu:5-18 */
void ctl_encode_tri_data(const float p[9], const float n[9], const float uv[6], uint32_t mat, ctl_tri_data* out); /* TriangleData.cu:8-65 */

/* ---- sample tables (Kernel/Sampler.h:22-85, Base/CudaRandom.h:108-291) --- */
/* Fill the 4096x30 1-D and 2-D tables of pass `pass` (0-based) of a fresh
 * tracer: XORWOW curand_init(1234, 7539414, 0), 4096*(30+60) draws per pass.
 * d1: n_seq*seq_len floats, d2: n_seq*seq_len*2 floats, element (seq,dim) at dim*n_seq+seq. */
int ctl_generate_sample_tables(uint32_t pass, float* d1, float* d2);
/* Passes first .. first+n-1 into n consecutive table sets (set k at d1 + k*4096*30, d2 + k*4096*60), produced concurrently on host threads: the start
 * state of every pass comes from a GF(2) jump-ahead of 4096*90 draws, each thread then runs the sequential generator.  Bit-identical to n calls above. */
int ctl_generate_sample_tables_n(uint32_t first_pass, int n, float* d1, float* d2);

/* ---- tracer context (Kernel/Tracer.h:67-294, Integrators/PathTracer.h:7-24) --- */

typedef struct ctl_ctx ctl_ctx;

const char* ctl_last_error(void);
/* == PathTracer ctor + TracerBase::Resize (Kernel/Tracer.h:102-109) */
ctl_ctx* ctl_create(int device, int width, int height);
void     ctl_destroy(ctl_ctx*);                          /* == TracerBase::~TracerBase (Kernel/Tracer.h:101) + Image::Free (Engine/Image.cpp:25) */
int      ctl_resize(ctl_ctx*, int width, int height);    /* == Tracer<true>::Resize (Kernel/Tracer.h:196-207): new PixelData / variance / queue storage, starts a new trace */
/* == m_sParameters: "MaxPathLength" (50), "RRStartDepth" (5), "Direct" (1),
 *    "Regularization" (0; 1 = PathTraceRegularization<DIRECT> instead of PathTrace<DIRECT>, Integrators/PathTracer.cu:115-170: all lights per vertex, no emitter
 *    MIS, one ray traced past the last vertex; needs MaxPathLength <= 255)  (Integrators/PathTracer.h:10-20);
 *    extras: "SortMode" (0 none, 1 material), "StageTimers" (0/1), "CaptureBounce" (0 = off),
 *    "DeviceSampleTables" (1 = tables generated by a CUDA kernel, bit-identical to the host XORWOW generator; 0 = generated on the
 *    host and copied H2D every pass like the reference's UpdateKernel), "TraversalKernel" (2 = persistent warps with shared-memory staging [default], 0 = persistent warps, 1 = ray-batch A/B baseline), "StagedThreads" / "StagedStackRows" /
 *    "StagedTreeletNodes" / "StagedResidentThreads" (launch shape of kernel 2; treelet nodes take effect at the next scene upload), "ShadeMode" (1 = one shade launch per
 *    material class [default], 0 = run-time BSDF dispatch), "StopZeroThroughput" (1 = a path of exactly zero throughput ends [default]; 0 = it is traced until Russian
 *    roulette, the reference's ray count), "TravTSteps", "FuseTraversal" (1 = shadow rays of bounce b and
 *    extension rays of bounce b+1 share one traversal launch),
 *    "TraversalBlocksPerSM", "TravThT/L/F", "TravThNExit", "TravChunk" (rays per queue claim, default 32), "TravDrainPrefetch" (tuning);
 *    frames (ctl_render_frame_tiled / ctl_comm_render_frame): "OverlapWavefronts" (0 / 1 [default: wavefronts of a frame on several streams when the frame has
 *    several] / 2 [also cut the batches of a one-wavefront frame]), "OverlapLanes" (1..8, default 4), "HandOver" (0 [default] / 1: a one-wavefront frame as two
 *    interleaved half-wavefronts whose traversal launches hand their unfinished rays over instead of draining; measured: correct, not faster -- DESIGN.md
 *    section 5), "HandOverDrain" (loop iterations a warp keeps going after the queue ran dry, default 16), "DeferStragglers" (0 [default] / 1: traversal launches move their unfinished rays into the
 *    wavefront's next launch, the paths lag up to "DeferMaxLag" = 1..3 bounces; measured: +2 % at 1/8 of the image, -2 % on the whole image), "ShadeConcurrent" (0 [default] / 1: the per-class
 *    shade launches of a bounce on their own streams; measured: -2 % on configs[2], a loss with lanes). */
int ctl_set_param_i(ctl_ctx*, const char* key, int value);   /* == TracerParameterCollection::setValue<int> (Kernel/TracerSettings.h:277-283) */
int ctl_get_param_i(ctl_ctx*, const char* key, int* value);  /* == getValue<int> (Kernel/TracerSettings.h:266-272) */
/* == UpdateKernel scene half (Kernel/TraceHelper.cu:182-217): host view copied to HBM */
int ctl_upload_scene(ctl_ctx*, const ctl_scene_view*);
/* Node-level half of ctl_upload_scene, for a view that differs from the uploaded one only above the meshes (instance transforms, scene-level BVH,
 * lights, box, epsilon, camera): what DynamicScene's Stream<T>::UpdateInvalidated re-uploads after SetNodeTransform.  Synchronises the stream. */
int ctl_update_scene_nodes(ctl_ctx*, const ctl_scene_view*);
/* == UpdateKernel sampler half / GenerateNewRandomSequences; host tables, async H2D */
int ctl_upload_samples(ctl_ctx*, const float* d1, const float* d2);
/* == __internal__IntersectBuffers (Kernel/TraceHelper.cu:736-746). Device pointers,
 *    n rays of ctl_traversal_ray -> n ctl_traversal_result. Asynchronous on `stream`
 *    (cudaStream_t, may be NULL = context stream); unlike the reference it does not
 *    synchronise -- call ctl_synchronize. */
int ctl_intersect(ctl_ctx*, int n, const void* d_rays, void* d_results, int any_hit, void* stream);
/* Host-buffer convenience: H2D, intersect, D2H, synchronous. */
int ctl_intersect_host(ctl_ctx*, int n, const ctl_traversal_ray* rays, ctl_traversal_result* results, int any_hit);
/* == traceRay (Kernel/TraceHelper.cu:174-180) batched: rays with tmin/tmax ignored
 *    (t in (rayEps, FLT_MAX)), full-precision barycentrics. Host buffers, synchronous.
 *    counts (may be NULL): [0]=inner nodes popped, [1]=triangle refs tested,
 *    [2]=instance leaves entered, summed over rays (instrumented build). */
int ctl_trace_rays_host(ctl_ctx*, int n, const ctl_traversal_ray* rays, ctl_trace_result* results, uint64_t counts[3]);
/* == Tracer<true>::DoPass (Kernel/Tracer.h:209-248) restricted to the pixel window
 *    [x0,x1) x [y0,y1) (full image: 0,0,w,h).  new_trace != 0 clears the accumulator
 *    and restarts the sample-table stream at pass 0.  Generates this pass's sample
 *    tables (unless tables were supplied by ctl_upload_samples since the last pass),
 *    then runs the wavefront stages.  Asynchronous; ctl_synchronize to wait. */
int ctl_render_pass(ctl_ctx*, int new_trace, int x0, int y0, int x1, int y1);
/* Interleaved-tile variant for multi-GPU: renders tiles (tile_w x tile_h) whose
 * index % n_parts == part. */
int ctl_render_pass_tiled(ctl_ctx*, int new_trace, int tile_w, int tile_h, int part, int n_parts);
/* n_passes consecutive DoPass calls fused into ONE wavefront (paths = pixels x passes; every path uses the sample tables
 * of its own pass, so each path is identical to the one the sequential passes would trace; only the order of the float
 * atomics into PixelData differs).  Keeps launches large when the image is split over many GPUs and amortises launch
 * overhead; path state is ~230 B x pixels x n_passes of HBM.  part=0, n_parts=1 renders the whole image. */
int ctl_render_passes_tiled(ctl_ctx*, int new_trace, int n_passes, int tile_w, int tile_h, int part, int n_parts);
/* One progressive FRAME (== StartNewTrace + spp DoPass calls, Kernel/Tracer.h:209-248) on the tiles of `part`, `batch` passes fused per wavefront
 * (spp % batch == 0).  With "OverlapWavefronts" = 1 (the default) a frame of SEVERAL wavefronts runs them on up to "OverlapLanes" (default 4) streams with
 * their own wavefront buffers, so that the draining end of one persistent traversal launch overlaps the head of another lane's (configs[4] at 1/8 of the
 * image per GPU: 457.9 -> 331.9 ms per frame); = 2 also cuts a frame of fewer wavefronts than lanes into smaller batches (measured on the 1 M-triangle
 * scene at 1/8 of the image: 37.6 ms per frame with or without -- what the overlap gains the extra launches lose, DESIGN.md section 5 -- hence not the
 * default; one-wavefront frames overlap with the NEXT frame instead: ctl_submit_frame_tiled below); = 0 never.  The paths traced are identical either way.
 * Asynchronous on the context's stream (the other streams are joined before the call returns work to it). */
int ctl_render_frame_tiled(ctl_ctx*, int spp, int batch, int tile_w, int tile_h, int part, int n_parts);
/* FRAMES IN FLIGHT -- a sequence of frames (== repeated StartNewTrace + spp DoPass calls, Kernel/Tracer.h:209-248; the reference finishes every pass with a
 * device synchronise, Kernel/TraceHelper.cu:744-745) as a pipeline.  ctl_submit_frame_tiled enqueues one whole frame -- accumulator clear, sample tables,
 * all wavefronts -- on a wavefront lane of its own (stream, wavefront buffers, PixelData accumulator, table sets); up to "FramesInFlight" (ctl_set_param_i,
 * 1..7, default 2) frames may be outstanding.  ctl_acquire_frame makes the context's stream wait for the OLDEST outstanding frame and makes that frame's
 * accumulator the context's accumulator (ctl_accum_device_ptr, ctl_resolve_*, ctl_apply_image_pipeline, ctl_read_accum then see it; it stays valid until
 * FramesInFlight further frames have been submitted); ctl_stats afterwards reports that frame's device time on its lane.  A lane starts after the work
 * the context's stream held when the frame was submitted.  The frames are the frames ctl_render_frame_tiled renders (same paths; only the order of the float
 * atomics differs); what the pipeline buys is that the drain of every persistent traversal launch -- the scene's longest rays, ~0.45 ms per launch whatever
 * its size -- is filled by the other frames' launches.  Other render calls fail while frames are outstanding.  Asynchronous. */
int ctl_submit_frame_tiled(ctl_ctx*, int spp, int batch, int tile_w, int tile_h, int part, int n_parts);
int ctl_acquire_frame(ctl_ctx*);
int ctl_frames_in_flight(ctl_ctx*);   /* frames submitted and not yet acquired */
/* == WavefrontPathTracer: Tracer<true>::DoPass + WavefrontPathTracer::DoRender (Kernel/Tracer.h:209-248,
 *    Integrators/PseudoRealtime/WavefrontPathTracer.cu:166-191) over a DoubleRayBuffer-shaped device queue (Kernel/DoubleRayBuffer.h):
 *    the reference's own wavefront integrator, second consumer of the intersect kernel.  Same parameters as the reference
 *    class ("Direct", "MaxPathLength", "RRStartDepth", WavefrontPathTracer.h:32-42).  One pass = one path per pixel; results equal the
 *    reference's algorithm run in the serial order of its queue atomics (see csrc/wavefront_pt.cuh).  Asynchronous.
 *    ctl_get_queue_sizes afterwards: ext[i] = primary rays intersected before iteration i, shadow[i] = secondary rays pushed by iteration i.
 *    "PassStride" / "PassPhase" (ctl_set_param_i, default 1 / 0): the k-th pass since the last new trace is pass PassPhase + k * PassStride
 *    of the frame (sample tables, sampler skip) -- several devices share a frame by pass index and sum their accumulators. */
int ctl_wavefront_pass(ctl_ctx*, int new_trace);
/* One progressive frame of the WavefrontPathTracer: a new trace of `spp` passes (== StartNewTrace + spp DoPass calls).  The passes are independent given
 * their index, so with "OverlapWavefronts" they run on up to "OverlapLanes" streams with their own queue buffers and the drain of one pass's traversal
 * launches is filled by another pass's; same passes, same paths, PixelData equal up to the order of the float atomics.  Asynchronous. */
int ctl_wavefront_frame(ctl_ctx*, int spp);
/* Device copy-back of sample-table set `table_set` (0 .. passes of the last batch - 1) for verification. */
int ctl_read_sample_tables(ctl_ctx*, int table_set, float* d1, float* d2);
/* == the ThrowCudaErrors(cudaDeviceSynchronize()) the reference ends every launch with (Kernel/TraceHelper.cu:745), on this context's stream only */
int ctl_synchronize(ctl_ctx*);
/* == Image accumulator: PixelData[w*h], reference layout (Engine/Image.h:10-29, getPixelData Image.h:64). */
int ctl_read_accum(ctl_ctx*, ctl_pixel_data* host_out);
/* == applyImagePipeline(tracer, img, 0, 0) (Kernel/ImagePipeline/ImagePipeline.cu:14-21, 54-63): the default resolve
 * PixelData -> toSpectrum(splatScale) -> sRGB -> RGBA8 (uchar4, a = 255).  Writes w*h*4 bytes to d_rgba8 (device,
 * asynchronous) and/or host_rgba8 (synchronous); either may be NULL.  First row of SURVEY 8f3 ("next"). */
int ctl_resolve_srgb8(ctl_ctx*, float splat_scale, void* d_rgba8, void* host_rgba8);
/* == applyImagePipeline(tracer, img, filter) (ImagePipeline.cu:70-74; the call the reference's example main makes with
 * BoxFilter(0.5, 0.5), main.cpp:172): CanonicalFilter reconstruction (Kernel/ImagePipeline/Filter/CanonicalFilter.cu:6-36) into the
 * RGBE stage, then gamma.  filter_type 0 = BoxFilter, 1 = GaussianFilter(alpha), 2 = TriangleFilter (SceneTypes/Filter.h). */
int ctl_resolve_filtered_srgb8(ctl_ctx*, float splat_scale, int filter_type, float x_width, float y_width, float alpha, void* d_rgba8, void* host_rgba8);
/* == applyImagePipeline(tracer, img, filter, process) in full (ImagePipeline.cu:54-84): optional CanonicalFilter (Box / Gaussian /
 * Triangle / Mitchell / LanczosSinc) into the RGBE stage, optional ToneMapPostProcess (Image::ComputeLuminanceInfo, Engine/Image.cu:88-173,
 * + Reinhard05Kernel, ToneMapPostProcess.cu:6-39) and the gamma stage.  lum_info (may be NULL, filled when tonemap != 0, synchronises):
 * [0] min [1] max [2] average luminance, [3] log-average luminance, [4] scale, [5] invWp2. */
int ctl_apply_image_pipeline(ctl_ctx*, float splat_scale, const ctl_image_pipeline*, void* d_rgba8, void* host_rgba8, float lum_info[6]);
/* filter_type 5 == NonLocalMeansFilter::Apply (Kernel/ImagePipeline/Filter/NonLocalMeansFilter.cu:184-228) in the filter slot of the pipeline: variance-guided
 * non-local means, 13x13 search window, 7x7 patches.  Needs "PixelVarianceBuffer"=1 during the passes.  Like the reference object, the context keeps the
 * weight buffer (169 floats per pixel) between calls and recomputes it when the pass count did not advance by exactly one since the last call, when it is a
 * multiple of UpdateWeightPeriodicity, or after a resize; otherwise the stored weights are applied to the new frame.  ctl_read_nlm_weights copies the
 * weights of the last call to the host in the reference's layout: w*h*169 floats, slot (yo + 6) * 13 + (xo + 6) of pixel y*w + x. */
int ctl_read_nlm_weights(ctl_ctx*, float* host_out);
/* == PixelVarianceBuffer (Kernel/PixelVarianceBuffer.h, .cu:10-36; owned by TracerBase, updated after every pass of a progressive tracer,
 * Kernel/Tracer.h:233-237): "PixelVarianceBuffer"=1 (ctl_set_param_i) makes every single-pass render call (ctl_render_pass with the full
 * window, ctl_wavefront_pass) run PixelVarianceInfo::updateMoments on all pixels after the pass and clear the buffer on a new trace.
 * ctl_read_variance copies the w*h records to the host. */
int ctl_read_variance(ctl_ctx*, ctl_pixel_variance_info* host_out);
/* Device pointer of the accumulator (7*w*h floats) for in-place NCCL reduce. */
void* ctl_accum_device_ptr(ctl_ctx*);
/* Use caller-owned device memory (7*w*h floats) as the accumulator (e.g. a torch tensor). */
int ctl_set_accum_device_ptr(ctl_ctx*, void* d_ptr);
/* == getRaysInLastPass / getLastTimeSpentRenderingSec (Kernel/Tracer.h:133-148);
 *    synchronises. rays = extension + shadow queries of the last pass. */
int ctl_stats(ctl_ctx*, uint64_t* rays_last_pass, float* seconds_last_pass, uint64_t* rays_total, uint32_t* passes_done);
/* Per-stage device time (ms) of the last pass: [0] generate [1] extension traversal
 * [2] shade [3] shadow traversal [4] accumulate/compact/sort; and launches. */
int ctl_stage_times(ctl_ctx*, float ms[5], uint32_t* n_launches);
/* Instrumented traversal of the NEXT pass: collects visit counts (slower).
 * counts[0..2] as in ctl_trace_rays_host, [3] = rays; for extension (0) / shadow (1). */
int ctl_set_instrumented(ctl_ctx*, int on);
int ctl_get_visit_counts(ctl_ctx*, uint64_t ext_counts[4], uint64_t shadow_counts[4]);
/* ctl_set_param_i(ctx, "CaptureBounce", b) makes every following pass keep a copy of the extension-ray
 * queue of bounce b (1-based; 0 = off). Fetch it (32-byte traversalRay records) after the pass; returns the
 * number of rays copied or -1.  Used for the ray-level micro-benchmark. */
int ctl_get_captured_rays(ctl_ctx*, ctl_traversal_ray* host_out, int capacity);
/* Extension / shadow queue sizes per bounce of the last pass (n entries each). */
int ctl_get_queue_sizes(ctl_ctx*, uint32_t* ext, uint32_t* shadow, int n);
/* The context's cudaStream_t (so callers can order their own work after the passes). */
void* ctl_stream(ctl_ctx*);
/* Run all following work of this context on a caller-owned cudaStream_t (NULL = back to the context's own
 * stream).  The reference is single-stream (default stream, Kernel/TraceHelper.cu:744); this lets a host
 * application order the passes with its own kernels / NCCL calls without extra synchronisation. */
int ctl_set_stream(ctl_ctx*, void* stream);


/* ---- multi-GPU: image tiles per rank + ONE NCCL reduce of the accumulator per frame (csrc/ctl_comm.cu) ---------------------------------
 * No reference counterpart: the reference is single-GPU (Kernel/TraceHelper.cu:744 synchronises one device; no NCCL / MPI in its tree).
 * Random numbers are a pure function of (pass, pixel index, dimension) (Kernel/Sampler_device.h:91-107), so disjoint tiles on different
 * devices trace exactly the paths one device would, and summing the zero-initialised accumulators is exact.  NCCL is loaded at run time
 * (dlopen of libnccl.so.2) and only by these calls. */
#define CTL_COMM_ID_BYTES 128 /* == NCCL_UNIQUE_ID_BYTES */
/* Rank 0 of a multi-process job: CTL_COMM_ID_BYTES bytes to hand to the other ranks (file, socket, MPI, environment ...). */
int ctl_comm_get_unique_id(void* id_out);
/* One process (or thread) per GPU: join the communicator described by `id` as `rank` of `n_ranks` (collective: every rank calls it). */
int ctl_comm_init_rank(ctl_ctx*, const void* id, int rank, int n_ranks);
/* One process driving n contexts on n distinct devices: contexts[i] becomes rank i.  Use the *_all collectives below with it. */
int ctl_comm_init_all(ctl_ctx* const* contexts, int n);
int ctl_comm_rank(const ctl_ctx*, int* rank, int* n_ranks);
/* The one collective of the path: ncclReduce(sum) of the PixelData accumulator (7*w*h floats) to `root`, in place, on the context's stream
 * (ordered after the frame's kernels); asynchronous.  Without a communicator (single GPU) it is a no-op. */
int ctl_comm_reduce_accum(ctl_ctx*, int root);
int ctl_comm_reduce_accum_all(ctl_ctx* const* contexts, int n, int root); /* the same for ctl_comm_init_all communicators (groups the per-device calls) */
/* Sum of `count` (<= 64) host counters over the ranks, e.g. the ray counts of a frame; synchronous. */
int ctl_comm_allreduce_u64(ctl_ctx*, uint64_t* host_inout, int count);
/* One progressive frame shared by the ranks: `spp` passes (`batch` fused per wavefront) on this rank's interleaved tile x tile tiles
 * (tile index % ranks == rank; tile <= 0 = 64), then ctl_comm_reduce_accum(root).  Asynchronous. */
int ctl_comm_render_frame(ctl_ctx*, int spp, int batch, int tile, int root);
/* The pipelined form (frames in flight, see ctl_submit_frame_tiled): this rank's tiles of one more frame on its own lane, and the frame's ncclReduce to `root`
 * on the context's communication stream (its own high-priority stream: every rank enqueues the reduces in submission order, and neither the rendering lanes nor
 * the context's stream wait for the other ranks).  ctl_acquire_frame returns the frames in order; on the root the acquired accumulator is the whole image. */
int ctl_comm_submit_frame(ctl_ctx*, int spp, int batch, int tile, int root);
int ctl_comm_submit_frame_all(ctl_ctx* const* contexts, int n, int spp, int batch, int tile, int root); /* for ctl_comm_init_all communicators */
int ctl_comm_destroy(ctl_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* CTL_B200_H */
