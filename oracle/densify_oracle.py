"""TEST INFRASTRUCTURE ONLY — sequential torch restatement of the reference's densification
(gaussian_splatting/scene/gaussian_model.py:258-402) on a plain dict of tensors + Adam state, step by step as the
reference performs it: densify_and_clone (append) -> densification_postfix (statistics reset) -> densify_and_split
(append two children per parent, prune the parents) -> opacity / size prune.  Pinned on tests/golden/ref_densify.npz,
the output of the reference's own GaussianModel class (tests/golden/make_densify_golden.py runs it in the build
container); the product's single-gather implementation is compared with both, row for row."""
import torch


def _rotmat(r):
    q = r / torch.sqrt((r * r).sum(dim=1))[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def _append(S, new):          # cat_tensors_to_optimizer + densification_postfix (:303-342)
    for k in NAMES:
        S["p"][k] = torch.cat((S["p"][k], new[k]), dim=0)
        S["m"][k] = torch.cat((S["m"][k], torch.zeros_like(new[k])), dim=0)
        S["v"][k] = torch.cat((S["v"][k], torch.zeros_like(new[k])), dim=0)
    n = S["p"]["xyz"].shape[0]
    S["accum"], S["denom"], S["max_radii2D"] = torch.zeros((n, 1)), torch.zeros((n, 1)), torch.zeros(n)


def _prune(S, mask):          # prune_points (:285-301)
    valid = ~mask
    for k in NAMES:
        for w in ("p", "m", "v"):
            S[w][k] = S[w][k][valid]
    S["accum"], S["denom"], S["max_radii2D"] = S["accum"][valid], S["denom"][valid], S["max_radii2D"][valid]


def densify_and_prune(S, max_grad, min_opacity, extent, max_screen_size, percent_dense, samples=None):
    """S = {"p": {name: tensor}, "m": {...}, "v": {...}, "accum": [N,1], "denom": [N,1], "max_radii2D": [N]} (CPU)."""
    grads = S["accum"] / S["denom"]
    grads[grads.isnan()] = 0.0
    scaling = lambda: torch.exp(S["p"]["scaling"])
    # densify_and_clone (:381-396)
    sel = (torch.norm(grads, dim=-1) >= max_grad) & (torch.max(scaling(), dim=1).values <= percent_dense * extent)
    _append(S, {k: S["p"][k][sel] for k in NAMES})
    # densify_and_split (:356-379), N = 2
    n = S["p"]["xyz"].shape[0]
    padded = torch.zeros(n)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = (padded >= max_grad) & (torch.max(scaling(), dim=1).values > percent_dense * extent)
    stds = scaling()[sel].repeat(2, 1)
    if samples is None:          # the goldens carry the reference's own draw (CPU and CUDA generators differ)
        samples = torch.normal(mean=torch.zeros((stds.size(0), 3)), std=stds)
    rots = _rotmat(S["p"]["rotation"][sel]).repeat(2, 1, 1)
    new = {"xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + S["p"]["xyz"][sel].repeat(2, 1),
           "scaling": torch.log(scaling()[sel].repeat(2, 1) / (0.8 * 2)), "rotation": S["p"]["rotation"][sel].repeat(2, 1),
           "f_dc": S["p"]["f_dc"][sel].repeat(2, 1, 1), "f_rest": S["p"]["f_rest"][sel].repeat(2, 1, 1),
           "opacity": S["p"]["opacity"][sel].repeat(2, 1)}
    _append(S, new)
    _prune(S, torch.cat((sel, torch.zeros(2 * int(sel.sum()), dtype=bool))))
    # final prune (:389-400)
    mask = (torch.sigmoid(S["p"]["opacity"]) < min_opacity).squeeze()
    if max_screen_size:
        mask = mask | (S["max_radii2D"] > max_screen_size) | (scaling().max(dim=1).values > 0.1 * extent)
    _prune(S, mask)
    return S
