"""PLY output / input of full-shape inference results, file-compatible with the reference's writer
(ref: u_net_arch/data_utils.py:13-21 get_dtype_dict, :36-50 read_ply_ls, :52-68 write_ply; called from
qualitative_inference_test.py:577-588 with ["vertex", ...]).  The reference goes through the `plyfile` package, which is
not a dependency here: the same files are produced with numpy alone — binary little-endian, one element per array,
float32 properties named as the reference names them (vertex/point: x y z; normal: nx ny nz; intensity: variation;
anything else: scalar_<name>), the element comment line `plyfile` emits."""
import numpy as np


def get_dtype_dict(name):
    if name in ("vertex", "point"):
        return {"names": ["x", "y", "z"], "formats": ["f4", "f4", "f4"]}
    if name == "intensity":
        return {"names": ["variation"], "formats": ["f4"]}
    if name == "normal":
        return {"names": ["nx", "ny", "nz"], "formats": ["f4", "f4", "f4"]}
    return {"names": ["scalar_{}".format(name)], "formats": ["f4"]}


def write_ply(filename, params_in_ls, params_names_ls):
    """Arrays (n_i,) or (n_i, k_i) -> one PLY element per array, named by params_names_ls (same call as the reference)."""
    header, blobs = ["ply", "format binary_little_endian 1.0"], []
    for param, name in zip(params_in_ls, params_names_ls):
        param = np.asarray(param)
        if param.ndim == 1:
            param = param[:, None]
        spec = get_dtype_dict(name)
        rec = np.zeros(param.shape[0], dtype=np.dtype({"names": spec["names"], "formats": ["<f4"] * len(spec["names"])}))
        for i, prop in enumerate(spec["names"]):
            rec[prop] = param[:, i]
        header.append("element {} {}".format(name, param.shape[0]))
        header.append("comment Generated with write_ply.py")
        header.extend("property float {}".format(prop) for prop in spec["names"])
        blobs.append(rec.tobytes())
    header.append("end_header")
    with open(filename, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        for blob in blobs:
            f.write(blob)


def read_ply_ls(filename, names):
    """-> {name: float32 array (n, k)} for the requested elements of a binary little-endian / ascii PLY file."""
    with open(filename, "rb") as f:
        assert f.readline().strip() == b"ply", "not a PLY file"
        fmt, elements = None, []
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                assert tok[1] != "list", "list properties are not used by this pipeline"
                elements[-1][2].append((tok[2], tok[1]))
        kinds = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "int": "i4", "int32": "i4", "uint": "u4",
                 "uint32": "u4", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1", "short": "i2", "int16": "i2",
                 "ushort": "u2", "uint16": "u2"}
        out = {}
        for name, count, props in elements:
            if fmt == "ascii":
                rows = np.array([f.readline().split() for _ in range(count)], dtype=np.float64).reshape(count, len(props))
                table = {p: rows[:, i] for i, (p, _) in enumerate(props)}
            else:
                order = "<" if fmt == "binary_little_endian" else ">"
                rec = np.frombuffer(f.read(count * sum(np.dtype(kinds[k]).itemsize for _, k in props)),
                                    dtype=np.dtype([(p, order + kinds[k]) for p, k in props]), count=count)
                table = {p: rec[p] for p, _ in props}
            if name in names:
                out[name] = np.stack([table[p] for p in get_dtype_dict(name)["names"]], 1).astype(np.float32)
    return out
