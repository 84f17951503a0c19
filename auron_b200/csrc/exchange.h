// exchange.h -- NCCL all-to-all-v repartition
#pragma once
#include "common.h"

namespace auron {

void nccl_get_unique_id(uint8_t out[128]);
void nccl_init(const uint8_t id[128], int rank, int world, int device);
void nccl_finalize();
int nccl_world();
int nccl_rank();
// `sorted` is partition-contiguous (rows of partition p = [part_row_off[p], part_row_off[p+1])); returns the rows of the
// partitions this rank owns (p * world / num_parts == rank), gathered from every rank.
BatchPtr nccl_exchange(Ctx& ctx, const Batch& sorted, const std::vector<int64_t>& part_row_off, int64_t num_parts, int64_t* bytes_sent);

}  // namespace auron
