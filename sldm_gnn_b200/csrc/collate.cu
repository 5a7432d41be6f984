// collate.cu -- device-side mini-batch assembly, the step BEFORE the SageBlock.
//
// The reference trains on torch_geometric.loader.DataLoader batches (main.py:166-167, src/utils.py:218-223): its
// dataset puts every graph on the device one by one (src/dataset.py:75-89, torch.load(map_location=device)), then PyG's
// collate (torch_geometric/data/collate.py, Batch.from_data_list, pinned 2.7.0) builds the block-diagonal batch:
//   * every tensor attribute is concatenated along dim 0 -- except attributes whose name contains "index"
//     (edge_index), which are concatenated along the last dim after adding the graph's node offset;
//   * `batch` [N] holds the graph id of every node, `ptr` [G+1] the node offsets.
// With G device-resident graphs that is ~2G+ small kernels (G offset adds, the cats, repeat_interleave); here it is one
// launch per attribute: a table of {source pointer, destination offset, size} is uploaded once per attribute and a
// CTA row copies each chunk.  Pure byte / integer work: bit-exact by construction.  Bound: HBM (bytes in + bytes out).
#include "common.cuh"
#include <algorithm>

namespace sldm {

// one entry per graph; a/b/c: see the kernels
struct ChunkRow { const void* src; int64_t a, b, c; };
static_assert(sizeof(ChunkRow) == 32, "ChunkRow is four 64-bit words on the host side too");

// out[a .. a + b) = src[0 .. b)   (bytes); 16-byte vectors when src, dst offset and size allow it
__global__ void __launch_bounds__(256)
k_concat_chunks(const ChunkRow* __restrict__ table, uint8_t* __restrict__ out) {
  const ChunkRow r = table[blockIdx.y];
  const uint8_t* __restrict__ src = static_cast<const uint8_t*>(r.src);
  uint8_t* __restrict__ dst = out + r.a;
  const int64_t bytes = r.b;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes) & 15u) == 0) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int64_t i = tid; i < bytes / 16; i += nth) d4[i] = __ldg(s4 + i);
  } else if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes) & 3u) == 0) {
    const int32_t* s1 = reinterpret_cast<const int32_t*>(src);
    int32_t* d1 = reinterpret_cast<int32_t*>(dst);
    for (int64_t i = tid; i < bytes / 4; i += nth) d1[i] = __ldg(s1 + i);
  } else {
    for (int64_t i = tid; i < bytes; i += nth) dst[i] = src[i];
  }
}

// edge_index of graph g: src = int64 [2, e_g] (row stride ld = c elements), a = edge offset, b = e_g, node offset in `noff`
// out is int64 [2, Etot]: out[0][a+k] = src[0][k] + noff, out[1][a+k] = src[1][k] + noff
__global__ void __launch_bounds__(256)
k_collate_edge_index(const ChunkRow* __restrict__ table, const int64_t* __restrict__ node_ptr, int64_t Etot,
                     int64_t* __restrict__ out) {
  const ChunkRow r = table[blockIdx.y];
  const int64_t* __restrict__ src = static_cast<const int64_t*>(r.src);
  const int64_t eoff = r.a, e = r.b, ld = r.c;
  const int64_t noff = node_ptr[blockIdx.y];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = tid; k < e; k += nth) {
    out[eoff + k] = src[k] + noff;
    out[Etot + eoff + k] = src[ld + k] + noff;
  }
}

// batch[i] = g for node_ptr[g] <= i < node_ptr[g+1]
__global__ void __launch_bounds__(256)
k_batch_from_ptr(const int64_t* __restrict__ node_ptr, int64_t* __restrict__ batch) {
  const int64_t beg = node_ptr[blockIdx.y], end = node_ptr[blockIdx.y + 1];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = beg + tid; i < end; i += nth) batch[i] = (int64_t)blockIdx.y;
}

static int grid_x(int64_t max_items) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(max_items, 256 * 4), 64));
}

}  // namespace sldm

using namespace sldm;

// table_dev: G rows of 4 x int64 {source device pointer, destination byte offset, byte count, unused}.
// max_chunk_bytes sizes the grid (largest chunk); chunks may be empty.
extern "C" int sldm_concat_chunks(const void* table_dev, int64_t G, int64_t max_chunk_bytes, void* out,
                                  sldm_stream_t stream) {
  SLDM_REQUIRE(G >= 0 && max_chunk_bytes >= 0, SLDM_EINVAL, "sldm_concat_chunks: negative size");
  if (G == 0 || max_chunk_bytes == 0) return SLDM_OK;
  SLDM_REQUIRE(table_dev != nullptr && out != nullptr, SLDM_EINVAL, "sldm_concat_chunks: NULL pointer");
  SLDM_REQUIRE(G <= 65535, SLDM_EUNSUPPORTED, "sldm_concat_chunks: more than 65535 chunks per call");
  dim3 grid(grid_x(max_chunk_bytes / 16 + 1), (unsigned)G);
  k_concat_chunks<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const ChunkRow*>(table_dev),
                                                                      static_cast<uint8_t*>(out));
  SLDM_LAUNCH_CHECK("k_concat_chunks");
  return SLDM_OK;
}

// table_dev: G rows {edge_index_g device pointer (int64 [2, e_g]), edge offset, e_g, row stride in elements};
// node_ptr_dev: int64 [G+1] node offsets.  out: int64 [2, Etot].  batch_out (may be NULL): int64 [N].
extern "C" int sldm_collate_graph_index(const void* table_dev, const int64_t* node_ptr_dev, int64_t G, int64_t Etot,
                                        int64_t max_edges, int64_t max_nodes, int64_t* edge_index_out,
                                        int64_t* batch_out, sldm_stream_t stream) {
  SLDM_REQUIRE(G >= 0 && Etot >= 0 && max_edges >= 0 && max_nodes >= 0, SLDM_EINVAL, "sldm_collate_graph_index: negative size");
  if (G == 0) return SLDM_OK;
  SLDM_REQUIRE(node_ptr_dev != nullptr, SLDM_EINVAL, "sldm_collate_graph_index: node_ptr is NULL");
  SLDM_REQUIRE(G <= 65535, SLDM_EUNSUPPORTED, "sldm_collate_graph_index: more than 65535 graphs per call");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (Etot > 0 && max_edges > 0) {
    SLDM_REQUIRE(table_dev != nullptr && edge_index_out != nullptr, SLDM_EINVAL, "sldm_collate_graph_index: NULL pointer");
    dim3 grid(grid_x(max_edges), (unsigned)G);
    k_collate_edge_index<<<grid, 256, 0, s>>>(static_cast<const ChunkRow*>(table_dev), node_ptr_dev, Etot, edge_index_out);
    SLDM_LAUNCH_CHECK("k_collate_edge_index");
  }
  if (batch_out != nullptr && max_nodes > 0) {
    dim3 grid(grid_x(max_nodes), (unsigned)G);
    k_batch_from_ptr<<<grid, 256, 0, s>>>(node_ptr_dev, batch_out);
    SLDM_LAUNCH_CHECK("k_batch_from_ptr");
  }
  return SLDM_OK;
}
