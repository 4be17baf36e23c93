"""Golden map file written and read back by the REFERENCE's own GaussianModel
(/root/reference/gaussian_splatting/scene/gaussian_model.py:177-208 construct_list_of_attributes / save_ply,
:215-256 load_ply), run on CPU in this container like make_densify_golden.py (module source executed with "cuda" -> "cpu").

The reference hands the structured array it builds (attribute order, transposes, normals, dtype 'f4') to the third-party
`plyfile` package, which is not in this image.  What `plyfile` does with it is fixed by the PLY format: a header that lists
the array's fields in order and, on a little-endian machine, the array's bytes (`PlyData([el]).write` defaults to
binary, native byte order).  `_PlyStandIn` below does exactly that and nothing else; every decision about WHAT is stored
where is the reference's code.  Reading goes through the same stand-in (`PlyData.read`, `elements[0][name]`,
`elements[0].properties[i].name` — the accessors load_ply uses).

    python tests/golden/make_ply_golden.py   ->  tests/golden/ref_map.ply  (bytes of the reference's save_ply)
                                                 tests/golden/ref_map.npz  (inputs + what the reference's load_ply returns)
"""
import os, sys, types
import numpy as np
import torch

REF = "/root/reference/gaussian_splatting"
HERE = os.path.dirname(os.path.abspath(__file__))


class _Property:
    def __init__(self, name):
        self.name = name


class _Element:
    def __init__(self, data, name):
        self.data, self.name = data, name
        self.properties = [_Property(n) for n in data.dtype.names]

    def __getitem__(self, key):
        return self.data[key]


class PlyElement:
    @staticmethod
    def describe(data, name):
        assert all(data.dtype[n] == np.dtype("f4") for n in data.dtype.names)
        return _Element(data, name)


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def write(self, path):
        el = self.elements[0]
        head = "ply\nformat binary_little_endian 1.0\nelement %s %d\n" % (el.name, el.data.shape[0])
        head += "".join("property float %s\n" % n for n in el.data.dtype.names) + "end_header\n"
        with open(path, "wb") as f:
            f.write(head.encode("ascii"))
            f.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline() == b"ply\n"
            assert f.readline() == b"format binary_little_endian 1.0\n"
            _, name, count = f.readline().decode().split()
            names = []
            while True:
                tok = f.readline().decode().split()
                if tok[0] == "end_header":
                    break
                assert tok[:2] == ["property", "float"]
                names.append(tok[2])
            data = np.frombuffer(f.read(), dtype=np.dtype([(n, "<f4") for n in names]), count=int(count))
        return PlyData([_Element(data, name)])


def load_patched(name, path):
    src = open(path).read().replace('"cuda"', '"cpu"').replace("'cuda'", "'cpu'").replace(".cuda()", ".cpu()")
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def main():
    sys.path.insert(0, REF)
    sys.modules["plyfile"] = types.SimpleNamespace(PlyData=PlyData, PlyElement=PlyElement)
    knn, knn_c = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
    knn_c.distCUDA2 = None
    sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn_c
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = pkg
    load_patched("utils.general_utils", os.path.join(REF, "utils", "general_utils.py"))
    gm = load_patched("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))

    g = torch.Generator().manual_seed(7)
    P, deg = 37, 3
    M = (deg + 1) ** 2
    r = lambda *s: torch.randn(*s, generator=g)
    params = dict(xyz=r(P, 3) * 2.0, f_dc=r(P, 1, 3) * 0.5, f_rest=r(P, M - 1, 3) * 0.05, opacity=r(P, 1) * 2.0,
                  scaling=r(P, 3) * 0.7 - 3.0, rotation=r(P, 4))
    model = gm.GaussianModel(deg)
    model._xyz, model._features_dc, model._features_rest = params["xyz"], params["f_dc"], params["f_rest"]
    model._opacity, model._scaling, model._rotation = params["opacity"], params["scaling"], params["rotation"]
    ply = os.path.join(HERE, "ref_map.ply")
    model.save_ply(ply)

    back = gm.GaussianModel(deg)
    back.load_ply(ply)
    out = {f"in_{k}": v.numpy() for k, v in params.items()}
    out.update(ld_xyz=back._xyz.detach().numpy(), ld_f_dc=back._features_dc.detach().numpy(),
               ld_f_rest=back._features_rest.detach().numpy(), ld_opacity=back._opacity.detach().numpy(),
               ld_scaling=back._scaling.detach().numpy(), ld_rotation=back._rotation.detach().numpy(),
               # the getters the renderer reads (gaussian_model.py:96-115)
               get_opacity=back.get_opacity.detach().numpy(), get_scaling=back.get_scaling.detach().numpy(),
               get_rotation=back.get_rotation.detach().numpy(), get_features=back.get_features.detach().numpy(),
               sh_degree=np.int32(back.active_sh_degree))
    np.savez_compressed(os.path.join(HERE, "ref_map.npz"), **out)
    print("wrote", ply, os.path.getsize(ply), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
