"""CPU oracle for mini-batch assembly -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/sage_oracle.py).

PARITY UNPINNED (torch-geometric 2.7.0 is not installable here): restates torch_geometric/data/collate.py as reached
through Batch.from_data_list by the reference's DataLoader (main.py:166-167; fields consumed at
src/models/grusage.py:153-173 and src/utils.py:219-223) with plain torch.cat:
  * `Data.__cat_dim__`: -1 for attributes whose name contains "index", else 0;
  * `Data.__inc__`: num_nodes (= x.size(0)) for attributes whose name contains "index", else 0;
  * `batch` = repeat_interleave(arange(G), num_nodes), `ptr` = cumsum of num_nodes with a leading 0.
"""
import torch


def collate_oracle(data_list):
    keys = [k for k in vars(data_list[0]) if not k.startswith("_")]
    nodes = [int(d.x.size(0)) for d in data_list]
    ptr = torch.zeros(len(data_list) + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(torch.tensor(nodes, dtype=torch.long), 0)
    out = {}
    for k in keys:
        vals = [getattr(d, k) for d in data_list]
        if not all(isinstance(v, torch.Tensor) for v in vals):
            out[k] = vals
        elif "index" in k:
            out[k] = torch.cat([v + int(ptr[g]) for g, v in enumerate(vals)], dim=-1)
        else:
            out[k] = torch.cat(vals, dim=0)
    out["batch"] = torch.repeat_interleave(torch.arange(len(data_list)), torch.tensor(nodes, dtype=torch.long))
    out["ptr"] = ptr
    out["num_graphs"] = len(data_list)
    return out
