"""BatchSOM — the assignment half of util/som.py:176-293 on the sm_100a kernels.

Only what the per-batch forward uses is implemented: the node buffer and query_topk / query.
Offline SOM training (batch_update / optimize, util/som.py:295-366) and the potential-field node
initialisation are preprocessing and out of scope.
"""
import torch

from . import ops


class BatchSOM():
    def __init__(self, rows=4, cols=4, dim=3, gpu_id=None, batch_size=10):
        self.rows = rows
        self.cols = cols
        self.dim = dim
        self.node_num = rows * cols
        self.gpu_id = gpu_id
        assert gpu_id is not None and gpu_id >= 0
        self.device = torch.device("cuda:%d" % gpu_id if torch.cuda.is_available() else "cpu")
        self.batch_size = batch_size
        # node: BxCx(rows*cols)
        self.node = torch.zeros(batch_size, dim, self.node_num, dtype=torch.float32,
                                device=self.device)
        self.node_idx_list = torch.arange(self.node_num, dtype=torch.int64, device=self.device)
        self.last_assignment = None  # dict from ops.som_assign for the most recent query

    def query_topk(self, x, k):
        """x [B,3,N] -> (mask [B,kN,M] int32, mask_row_max [B,M] int32, min_idx [B,kN] int64),
        util/som.py:237-269. Slot order: ascending distance (the reference's is unspecified)."""
        M = self.rows * self.cols
        node = self.node
        if node.shape[0] != x.shape[0]:
            node = node.expand(x.shape[0], node.shape[1], node.shape[2])
        mask, row_max, idx64, idx32 = ops.som_query_topk(x.detach().contiguous(), node.contiguous(), k)
        self.last_assignment = dict(min_idx_i32=idx32, min_idx_i64=idx64, row_max=row_max,
                                    count=None, cluster_mean=None)
        return mask, row_max, idx64

    def query(self, x):
        """k=1 variant (util/som.py:271-293): (mask [B,N,M] float, mask_row_max [B,M] float)."""
        M = self.rows * self.cols
        a = ops.som_assign(x.detach().contiguous(), self.node.contiguous(), 1, want_stats=False)
        self.last_assignment = a
        return ops.som_mask(a["min_idx_i32"], M).float(), a["row_max"].float()
