"""Mask helpers and host-side index builders (reference: pretrain_src/model/ops.py:25-68 and the Python
loops of vilmodel.py:632-666 / pretrain_cmt.py:339-356, turned into index tensors built once per batch)."""
import torch

NEG_MASK = -10000.0


def gen_seq_masks(seq_lens, max_len=None):
    """(N,) lengths -> (N, max_len) bool (ops.py:36-44)."""
    if max_len is None:
        max_len = int(seq_lens.max())
    return torch.arange(max_len, device=seq_lens.device)[None, :] < seq_lens[:, None]


def neg_key_mask(masks, neg=NEG_MASK):
    """bool (N,L) -> additive fp32 (N,L): 0 where valid, `neg` where padded (extend_neg_masks, ops.py:25-34,
    without the broadcast dims -- the softmax kernel broadcasts over heads and queries)."""
    return (1.0 - masks.to(torch.float32)) * neg


def inf_key_mask(masks):
    """bool (N,L) valid -> additive fp32 with -inf on padding (nn.MultiheadAttention key_padding_mask)."""
    return torch.zeros(masks.shape, dtype=torch.float32, device=masks.device).masked_fill(~masks, float("-inf"))


def build_gmap_segments(traj_step_lens, traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids, G, Vtot, device):
    """CSR lists for the topological-map node features (vilmodel.py:632-666).

    Node j of sample i (row i*G + j) is
      - the mean over the valid tokens of the panorama of a visited viewpoint, or
      - the mean of the candidate-view tokens (over all panoramas) that point at an unvisited viewpoint;
    row j = 0 ([stop]) and padded rows are empty segments (zeros).  Token (pano p, view v) is row p*Vtot+v of
    the flattened panorama embeddings.  `traj_vp_lens` is a host list of per-panorama token counts.
    """
    seg_off, idx, w = [0], [], []
    p0 = 0
    for i, P in enumerate(traj_step_lens):
        visited_at = {}
        cand_tok = {}
        for t in range(P):
            visited_at[traj_vpids[i][t]] = p0 + t
            for j, vp in enumerate(traj_cand_vpids[i][t]):
                if vp not in visited_at:
                    cand_tok.setdefault(vp, []).append((p0 + t) * Vtot + j)
        nodes = gmap_vpids[i]
        for j in range(G):
            if 0 < j < len(nodes):
                vp = nodes[j]
                if vp in visited_at:
                    pano = visited_at[vp]
                    n = int(traj_vp_lens[pano])
                    idx.extend(range(pano * Vtot, pano * Vtot + n))
                    w.extend([1.0 / n] * n)
                else:
                    toks = cand_tok[vp]
                    idx.extend(toks)
                    w.extend([1.0 / len(toks)] * len(toks))
            seg_off.append(len(idx))
        p0 += P
    return (torch.tensor(seg_off, dtype=torch.int32, device=device), torch.tensor(idx, dtype=torch.int32, device=device),
            torch.tensor(w, dtype=torch.float32, device=device))


def build_sap_fusion(gmap_vpids, gmap_visited_masks_host, last_cand_vpids, G, Kc, device):
    """(B, G, Kc) 0/1 fp32 matrix F with fused[b,j] = global[b,j] + sum_k F[b,j,k] * local[b,k]
    (pretrain_cmt.py:339-356): slot 0 takes local[0]; an unvisited node takes the local logit of the
    candidate view that reaches it, otherwise the sum over candidates that lead back to visited nodes."""
    B = len(gmap_vpids)
    F = torch.zeros(B, G, Kc, dtype=torch.float32)
    for i in range(B):
        visited = {vp for vp, m in zip(gmap_vpids[i], gmap_visited_masks_host[i]) if m}
        F[i, 0, 0] = 1.0
        direct, back = {}, []
        for j, vp in enumerate(last_cand_vpids[i]):
            if vp in visited:
                back.append(j + 1)
            else:
                direct[vp] = j + 1
        for j, vp in enumerate(gmap_vpids[i]):
            if j > 0 and vp not in visited:
                if vp in direct:
                    F[i, j, direct[vp]] = 1.0
                else:
                    for k in back:
                        F[i, j, k] = 1.0
    return F.to(device)
