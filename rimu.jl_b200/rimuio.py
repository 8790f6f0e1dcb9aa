"""RimuIO: `save_state` / `load_state` of walker vectors as Arrow files (host mirror of RimuIO/RimuIO.jl:92-195 and the
address serialisation of RimuIO/arrowtypes.jl:8-161).  Pure host code over the vector's downloaded (key, value) arrays.

File layout written by the reference (`Arrow.write(filename, Tables.table(vector); compress=:zstd, metadata)`):
  * Arrow IPC *file*, zstd-compressed buffers, two columns `key`, `value`;
  * schema metadata: every keyword as a string, plus `RIMU_PACKAGE_VERSION`;
  * `value`: Float64 / Int64;
  * `key`: the address's chunks as the Arrow image of an `NTuple{N,T}` (fixed-size list of N unsigned integers, chunk 1 =
    MOST significant, bitstring.jl:72-75) tagged with the ArrowTypes extension metadata
      BoseFS   -> name "Rimu.BoseFS.BitString",  metadata "N.M.B"   (B = N+M-1 bits)
      FermiFS  -> name "Rimu.FermiFS.BitString", metadata "N.M.B"   (B = M)
      CompositeFS -> a struct of the components' lists (fields "1", "2", ...), name "Rimu.CompositeFS",
                     metadata "name:meta;name:meta"
    with the chunk type of `num_chunks` (bitstring.jl:6-18): UInt8/16/32 for B <= 8/16/32, else ceil(B/64) x UInt64.
Differences to note: the reference stores sparse boson addresses (`SortedParticleList`, chosen when the dense form needs more
words than the particle list, bosefs.jl:85-97) under a different extension name; this module always writes the dense
`BitString` form (the only one the device path has) and refuses to read the sparse one.  Interoperability with Arrow.jl could
not be executed here (no Julia): the layout follows the reference's source; the tests check it field by field.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .addresses import AddressType

RIMU_PACKAGE_VERSION = "0.14.0"  # the reference version whose file layout is restated here
EXT_NAME, EXT_META = b"ARROW:extension:name", b"ARROW:extension:metadata"
_NAMES = {_lib.ADDR_BOSE: "Rimu.BoseFS.BitString", _lib.ADDR_FERMI: "Rimu.FermiFS.BitString"}


def chunk_layout(bits: int):
    """num_chunks(Val(B)) (bitstring.jl:6-18) -> (number of chunks, numpy dtype of a chunk)."""
    if bits <= 0:
        raise ValueError("`B` must be positive!")
    if bits <= 8:
        return 1, np.uint8
    if bits <= 16:
        return 1, np.uint16
    if bits <= 32:
        return 1, np.uint32
    return (bits - 1) // 64 + 1, np.uint64


def _to_int(words) -> int:
    x = 0
    for j, w in enumerate(words):
        x |= int(w) << (64 * j)
    return x


def _chunks_of(x: int, bits: int):
    """most-significant chunk first"""
    n, dt = chunk_layout(bits)
    if n == 1:
        return [x & ((1 << bits) - 1)]
    return [(x >> (64 * (n - 1 - i))) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def _components(at: AddressType):
    """[(extension name, N, M, B, shift in the packed device key)] per component"""
    M = at.num_modes
    if at.kind == _lib.ADDR_BOSE:
        return [(_NAMES[_lib.ADDR_BOSE], at.num_particles[0], M, at.num_particles[0] + M - 1, 0)]
    if at.kind == _lib.ADDR_COMPOSITE:  # general CompositeFS: the components side by side from the low bits
        out, shift = [], 0
        for k, n, b in zip(at.comp_kinds, at.num_particles, at.comp_bits):
            out.append((_NAMES[k], n, M, b, shift))
            shift += b
        return out
    return [(_NAMES[_lib.ADDR_FERMI], n, M, M, c * M) for c, n in enumerate(at.num_particles)]


def _key_arrays(keys, at: AddressType):
    """device keys (n, W) uint64 -> one (n, nchunks) array per component"""
    keys = np.asarray(keys, dtype=np.uint64).reshape(-1, at.words)
    out = []
    for name, N, M, B, shift in _components(at):
        nch, dt = chunk_layout(B)
        arr = np.zeros((len(keys), nch), dtype=dt)
        if at.words == 1 and nch == 1:  # vectorised common case
            arr[:, 0] = ((keys[:, 0] >> np.uint64(shift)) & np.uint64((1 << B) - 1)).astype(dt)
        else:
            for r in range(len(keys)):
                x = (_to_int(keys[r]) >> shift) & ((1 << B) - 1)
                arr[r, :] = _chunks_of(x, B)
        out.append(arr)
    return out


def _state_batch(keys, values, address_type: AddressType, metadata=None):
    """(schema, record batch) of a (key, value) table in the reference's layout"""
    import pyarrow as pa
    values = np.ascontiguousarray(values)
    if values.dtype not in (np.float64, np.int64):
        raise TypeError("values must be Float64 or Int64")
    comps = _components(address_type)
    lists = []
    for arr in _key_arrays(keys, address_type):
        flat = pa.array(arr.reshape(-1), type=pa.from_numpy_dtype(arr.dtype))
        lists.append(pa.FixedSizeListArray.from_arrays(flat, arr.shape[1]))
    if len(comps) == 1:
        key_arr = lists[0]
        name, meta = comps[0][0], f"{comps[0][1]}.{comps[0][2]}.{comps[0][3]}"
    else:  # CompositeFS: tuple of tuples = struct with fields "1", "2", ... (arrowtypes.jl:127-131)
        key_arr = pa.StructArray.from_arrays(lists, names=[str(i + 1) for i in range(len(lists))])
        name = "Rimu.CompositeFS"
        meta = ";".join(f"{c[0]}:{c[1]}.{c[2]}.{c[3]}" for c in comps)
    key_field = pa.field("key", key_arr.type, nullable=False, metadata={EXT_NAME: name.encode(), EXT_META: meta.encode()})
    val_field = pa.field("value", pa.from_numpy_dtype(values.dtype), nullable=False)
    md = {"RIMU_PACKAGE_VERSION": RIMU_PACKAGE_VERSION}
    md.update({str(k): _julia_string(v) for k, v in (metadata or {}).items()})
    schema = pa.schema([key_field, val_field], metadata={k.encode(): v.encode() for k, v in md.items()})
    return schema, pa.record_batch([key_arr, pa.array(values)], schema=schema)


def write_state_file(filename, keys, values, address_type: AddressType, metadata=None):
    """Arrow.write(filename, (key=..., value=...); compress=:zstd, metadata) for device-layout keys."""
    import pyarrow as pa
    schema, batch = _state_batch(keys, values, address_type, metadata)
    with pa.OSFile(str(filename), "wb") as sink:
        with pa.ipc.new_file(sink, schema, options=pa.ipc.IpcWriteOptions(compression="zstd")) as writer:
            writer.write_batch(batch)


_EOS = b"\xff\xff\xff\xff\x00\x00\x00\x00"  # end-of-stream marker of the Arrow IPC streaming format


def _stream_image(schema, batch):
    """bytes of an Arrow IPC *stream* holding one record batch -> (schema message, record-batch message(s))"""
    import pyarrow as pa
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, schema, options=pa.ipc.IpcWriteOptions(compression="zstd")) as writer:
        writer.write_batch(batch)
    raw = sink.getvalue().to_pybytes()
    assert raw[:4] == b"\xff\xff\xff\xff" and raw.endswith(_EOS)
    schema_end = 8 + int.from_bytes(raw[4:8], "little")  # a schema message has no body
    return raw[:schema_end], raw[schema_end:-len(_EOS)]


def write_state_stream(filename, keys, values, address_type: AddressType, metadata=None):
    """rank 0 of a multi-rank save: `Arrow.write(...; file=false)` (RimuIO.jl:113-117) -- the IPC STREAM format, which
    later ranks can extend"""
    schema, batch = _state_batch(keys, values, address_type, metadata)
    head, body = _stream_image(schema, batch)
    with open(str(filename), "wb") as f:
        f.write(head + body + _EOS)


def append_state_stream(filename, keys, values, address_type: AddressType):
    """`Arrow.append(filename, table)` (RimuIO.jl:122-126): one more record batch at the end of an IPC stream file"""
    schema, batch = _state_batch(keys, values, address_type, None)
    _, body = _stream_image(schema, batch)
    with open(str(filename), "r+b") as f:
        f.seek(-len(_EOS), 2)
        if f.read(len(_EOS)) != _EOS:
            raise ValueError(f"`{filename}` is not an Arrow IPC stream")
        f.seek(-len(_EOS), 2)
        f.truncate()
        f.write(body + _EOS)


def _julia_string(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    return str(v)


def _parse_meta(k, v):
    """RimuIO.jl:166-178: Int, then Float64, then ComplexF64, then Bool, else the string."""
    if k == "RIMU_PACKAGE_VERSION":
        return v
    try:
        return int(v)
    except ValueError:
        pass
    try:
        return float(v)
    except ValueError:
        pass
    try:
        return complex(v.replace("im", "j").replace(" ", ""))
    except ValueError:
        pass
    if v in ("true", "false"):
        return v == "true"
    return v


def _decode_component(name, meta):
    if name not in _NAMES.values():
        raise ValueError(f"address storage `{name}` has no device layout (only dense BitString addresses are supported)")
    N, M, B = (int(t) for t in meta.split("."))
    return (_lib.ADDR_BOSE if name == _NAMES[_lib.ADDR_BOSE] else _lib.ADDR_FERMI), N, M, B


def read_state_file(filename):
    """-> (keys (n, W) uint64 in the device layout, values, AddressType, metadata dict)"""
    import pyarrow as pa
    with pa.OSFile(str(filename), "rb") as src:
        try:
            tbl = pa.ipc.open_file(src).read_all()
        except pa.ArrowInvalid:  # the streaming format a multi-rank save writes (Arrow.Table reads both)
            src.seek(0)
            tbl = pa.ipc.open_stream(src).read_all()
    if tbl.schema.names != ["key", "value"]:
        raise ValueError(f"`{filename}` is not a valid Rimu state file")  # ArgumentError (RimuIO.jl:150-152)
    kf = tbl.schema.field("key")
    fmeta = kf.metadata or {}
    name, meta = fmeta.get(EXT_NAME, b"").decode(), fmeta.get(EXT_META, b"").decode()
    col = tbl.column("key").combine_chunks()
    if name == "Rimu.CompositeFS":
        parts = [m.split(":") for m in meta.split(";")]
        comps = [_decode_component(n, mm) for n, mm in parts]
        M = comps[0][2]
        if len(comps) == 2 and all(c[0] == _lib.ADDR_FERMI for c in comps) and M <= 32:
            at = AddressType(_lib.ADDR_FERMI2C, tuple(c[1] for c in comps), M)
        else:
            if not 2 <= len(comps) <= _lib.MAX_COMPONENTS or sum(c[3] for c in comps) > 127:
                raise ValueError("this CompositeFS has no device layout (2..4 components, at most 127 bits)")
            at = AddressType(_lib.ADDR_COMPOSITE, tuple(c[1] for c in comps), M, tuple(c[0] for c in comps))
        lists = [col.field(i) for i in range(len(comps))]
        bits = [c[3] for c in comps]
        shifts = [sum(bits[:i]) for i in range(len(comps))]
    else:
        kind, N, M, B = _decode_component(name, meta)
        at = AddressType(kind, (N,), M)
        lists, shifts, bits = [col], [0], [B]
    n = len(col)
    keys = np.zeros((n, at.words), dtype=np.uint64)
    big = [0] * n if at.words > 1 else None
    for lst, shift, B in zip(lists, shifts, bits):
        nch, dt = chunk_layout(B)
        if lst.type.list_size != nch:
            raise ValueError("chunk count does not match the declared number of bits")
        flat = lst.flatten().to_numpy(zero_copy_only=False).reshape(n, nch)
        if at.words == 1:
            keys[:, 0] |= flat[:, 0].astype(np.uint64) << np.uint64(shift)
        else:
            for r in range(n):
                x = 0
                for i in range(nch):  # chunk 0 is the most significant
                    x = (x << 64) | int(flat[r, i]) if nch > 1 else int(flat[r, i])
                big[r] |= x << shift
    if at.words > 1:
        for r in range(n):
            for j in range(at.words):
                keys[r, j] = (big[r] >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    values = tbl.column("value").combine_chunks().to_numpy(zero_copy_only=False)
    smeta = tbl.schema.metadata or {}
    metadata = {k.decode(): _parse_meta(k.decode(), v.decode()) for k, v in smeta.items()}
    return keys, values, at, metadata


# --------------------------------------------------------------------------- vector-level API
def save_state(filename, vector, **kwargs):
    """save_state(filename, vector; kwargs...) (RimuIO.jl:92-105).  One process; with several ranks every rank holds only its
    share of the vector (the reference appends the ranks' record batches to one file, RimuIO.jl:107-135)."""
    keys, vals = vector.download()
    ctx = vector.ctx
    if ctx.nranks == 1:
        write_state_file(filename, keys, vals, vector.address_type, kwargs)
        return
    # _save_state_mpi (RimuIO.jl:107-135): rank 0 creates the file (streaming format + metadata), then the other ranks append
    # their record batch one after the other, in rank order.  The barrier is the library's own all-reduce.
    if ctx.rank == 0:
        write_state_stream(filename, keys, vals, vector.address_type, kwargs)
    for r in range(1, ctx.nranks):
        ctx.allreduce([0.0])
        if ctx.rank == r:
            append_state_stream(filename, keys, vals, vector.address_type)
    ctx.allreduce([0.0])


def load_state(filename, style=None, ctx=None, **vector_kwargs):
    """load_state(filename; kwargs...) -> (vector, metadata) (RimuIO.jl:137-184): the style defaults to
    IsDynamicSemistochastic for Float64 values and to the integer style for Int64 values."""
    from .dictvectors import GPUDVec
    from .stochasticstyles import IsDynamicSemistochastic, IsStochasticInteger
    keys, vals, at, metadata = read_state_file(filename)
    if style is None:
        style = IsDynamicSemistochastic() if vals.dtype == np.float64 else IsStochasticInteger()
    v = GPUDVec(style=style, address_type=at, capacity=max(len(vals), 256), ctx=ctx, **vector_kwargs)
    nz = vals != 0
    if v.ctx.nranks > 1:
        # every rank reads the whole file and keeps the addresses it owns: upload() filters by owner rank (a plain
        # assign() would leave every rank with the full vector and multiply the population by the number of ranks)
        v.upload(keys[nz], vals[nz].astype(v.dtype))
    else:
        v.assign(keys[nz], vals[nz].astype(v.dtype))
    return v, metadata
