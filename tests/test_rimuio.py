"""RimuIO state files (RimuIO/RimuIO.jl:92-195, RimuIO/arrowtypes.jl:8-161): the Arrow layout is checked field by field
against what the reference's ArrowTypes definitions prescribe, and keys/values/metadata round-trip.  Host only (pyarrow)."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as orc


def _read_raw(path):
    with pa.OSFile(str(path), "rb") as f:
        return pa.ipc.open_file(f).read_all()


def test_chunk_layout_follows_num_chunks(built):
    from rimu_b200.rimuio import chunk_layout
    # bitstring.jl:6-18
    assert chunk_layout(8) == (1, np.uint8) and chunk_layout(9) == (1, np.uint16) and chunk_layout(19) == (1, np.uint32)
    assert chunk_layout(33) == (1, np.uint64) and chunk_layout(64) == (1, np.uint64) and chunk_layout(65) == (2, np.uint64)
    assert chunk_layout(127) == (2, np.uint64)
    with pytest.raises(ValueError):
        chunk_layout(0)


def test_bose_state_file_layout_and_roundtrip(built, tmp_path):
    import rimu_b200 as R
    from rimu_b200 import rimuio
    addrs = [R.BoseFS((0, 0, 3, 0, 2)), R.BoseFS((1, 1, 1, 1, 1)), R.BoseFS((5, 0, 0, 0, 0))]  # BoseFS{5,5}: 9 bits -> one UInt16 chunk
    at = addrs[0].address_type
    keys = np.array([a.key() for a in addrs], dtype=np.uint64).reshape(-1, 1)
    vals = np.array([1.5, -2.0, 0.25])
    path = tmp_path / "state.arrow"
    rimuio.write_state_file(path, keys, vals, at, dict(shift=-4.5, step=17, style="IsDynamicSemistochastic", flag=True))
    tbl = _read_raw(path)
    assert tbl.schema.names == ["key", "value"]                                  # RimuIO.jl:150
    kf = tbl.schema.field("key")
    assert kf.metadata[b"ARROW:extension:name"] == b"Rimu.BoseFS.BitString"       # arrowtypes.jl:61-62
    assert kf.metadata[b"ARROW:extension:metadata"] == b"5.5.9"                   # N.M.B, arrowtypes.jl:52-54,15
    assert pa.types.is_fixed_size_list(kf.type) and kf.type.list_size == 1 and kf.type.value_type == pa.uint16()
    assert tbl.schema.field("value").type == pa.float64()
    assert tbl.schema.metadata[b"RIMU_PACKAGE_VERSION"] == b"0.14.0" and tbl.schema.metadata[b"flag"] == b"true"
    # the stored chunk is the bit pattern of bitstring.jl:464-472 (mode 1 in the low bits: n ones then a 0)
    raw = tbl.column("key").combine_chunks().flatten().to_numpy()
    onr0 = (0, 0, 3, 0, 2)
    bits, pos = 0, 0
    for n in onr0:
        bits |= ((1 << n) - 1) << pos
        pos += n + 1
    assert int(raw[0]) == bits
    k2, v2, at2, meta = rimuio.read_state_file(path)
    assert at2 == at and np.array_equal(k2, keys) and np.array_equal(v2, vals)
    assert meta["shift"] == -4.5 and meta["step"] == 17 and meta["style"] == "IsDynamicSemistochastic" and meta["flag"] is True
    assert meta["RIMU_PACKAGE_VERSION"] == "0.14.0"


def test_two_word_and_fermion_state_files(built, tmp_path):
    import rimu_b200 as R
    from rimu_b200 import rimuio, _lib
    # BoseFS{64,64}: 127 bits -> two UInt64 chunks, chunk 1 MOST significant (bitstring.jl:72-75)
    oh = orc.OracleHam("HubbardRealSpace", "bose", tuple([1] * 64), u=1.0, t=1.0, dims=(4, 4, 4))
    k0 = np.array(oh.start_key, dtype=np.uint64).reshape(1, 2)
    nb = [oh.get_offdiagonal(tuple(int(x) for x in k0[0]), i)[0] for i in (1, 2, 200)]
    keys = np.concatenate([k0, np.array(nb, dtype=np.uint64).reshape(-1, 2)])
    vals = np.array([3, -1, 2, 7], dtype=np.int64)
    at = R.AddressType(_lib.ADDR_BOSE, (64,), 64)
    path = tmp_path / "w2.arrow"
    rimuio.write_state_file(path, keys, vals, at)
    tbl = _read_raw(path)
    kf = tbl.schema.field("key")
    assert kf.metadata[b"ARROW:extension:metadata"] == b"64.64.127" and kf.type.list_size == 2 and kf.type.value_type == pa.uint64()
    raw = tbl.column("key").combine_chunks().flatten().to_numpy().reshape(-1, 2)
    assert np.array_equal(raw[:, 0], keys[:, 1]) and np.array_equal(raw[:, 1], keys[:, 0])  # [high word, low word]
    assert tbl.schema.field("value").type == pa.int64()
    k2, v2, at2, _ = rimuio.read_state_file(path)
    assert at2 == at and np.array_equal(k2, keys) and np.array_equal(v2, vals)
    # FermiFS2C = CompositeFS of two FermiFS{N,M}: struct of two chunk lists (arrowtypes.jl:113-161)
    a = R.FermiFS2C((1, 0, 1, 0, 0, 1), (0, 1, 0, 0, 1, 0))
    b = R.FermiFS2C((0, 1, 1, 0, 0, 1), (1, 0, 0, 0, 1, 0))
    fk = np.array([a.key(), b.key()], dtype=np.uint64).reshape(-1, 1)
    fv = np.array([0.5, -0.75])
    p2 = tmp_path / "f2c.arrow"
    rimuio.write_state_file(p2, fk, fv, a.address_type)
    t2 = _read_raw(p2)
    kf = t2.schema.field("key")
    assert kf.metadata[b"ARROW:extension:name"] == b"Rimu.CompositeFS"
    assert kf.metadata[b"ARROW:extension:metadata"] == b"Rimu.FermiFS.BitString:3.6.6;Rimu.FermiFS.BitString:2.6.6"
    assert pa.types.is_struct(kf.type) and [kf.type.field(i).name for i in range(2)] == ["1", "2"]
    assert kf.type.field(0).type.value_type == pa.uint8()  # 6 bits -> UInt8 chunk
    col = t2.column("key").combine_chunks()
    up = col.field(0).flatten().to_numpy()
    assert int(up[0]) == 0b100101  # mode m <-> bit m-1 (bitstring.jl:713-723): modes 1, 3, 6
    k3, v3, at3, _ = rimuio.read_state_file(p2)
    assert at3 == a.address_type and np.array_equal(k3, fk) and np.array_equal(v3, fv)
    # a file that is not a state file
    bad = tmp_path / "bad.arrow"
    with pa.OSFile(str(bad), "wb") as f:
        with pa.ipc.new_file(f, pa.schema([("x", pa.int64())])) as w:
            w.write_batch(pa.record_batch([pa.array([1, 2])], names=["x"]))
    with pytest.raises(ValueError):
        rimuio.read_state_file(bad)


def test_save_and_load_state_wrappers(built, tmp_path, monkeypatch):
    """save_state / load_state (RimuIO.jl:92-105,137-184) over a duck-typed vector: what they do to a device vector is
    download() and assign(), both covered by the GPU parity tests; here the glue (metadata, default styles, zero
    filtering, multi-rank refusal) runs without a device."""
    import rimu_b200 as R
    from rimu_b200 import rimuio, dictvectors

    class Ctx:
        nranks = 1

    class FakeVec:
        def __init__(self, keys=None, vals=None, style=None, address_type=None, capacity=0, ctx=None, **kw):
            self.keys, self.vals, self.style, self.address_type, self.ctx, self.kw = keys, vals, style, address_type, ctx or Ctx(), kw
            self.dtype = np.int64 if style is not None and style.val_type == R._lib.VAL_I64 else np.float64

        def download(self):
            return self.keys, self.vals

        def assign(self, keys, vals):
            self.keys, self.vals = np.array(keys), np.array(vals)

    addrs = [R.BoseFS((2, 0, 1)), R.BoseFS((1, 1, 1)), R.BoseFS((0, 3, 0))]
    keys = np.array([a.key() for a in addrs], dtype=np.uint64).reshape(-1, 1)
    src = FakeVec(keys, np.array([4, 0, -2], dtype=np.int64), address_type=addrs[0].address_type)
    path = tmp_path / "v.arrow"
    R.save_state(path, src, shift=1.25, laststep=100)
    monkeypatch.setattr(dictvectors, "GPUDVec", FakeVec)
    v, meta = R.load_state(path, initiator=R.Initiator(2.0))
    assert isinstance(v.style, R.IsStochasticInteger) and v.address_type == addrs[0].address_type
    assert np.array_equal(v.keys, keys[[0, 2]]) and np.array_equal(v.vals, [4, -2])   # zeros are never stored
    assert v.kw == {"initiator": R.Initiator(2.0)}                                     # kwargs reach the vector constructor
    assert meta["shift"] == 1.25 and meta["laststep"] == 100
    src2 = FakeVec(keys, np.array([0.5, 1.5, -2.0]), address_type=addrs[0].address_type)
    R.save_state(path, src2)
    v2, _ = R.load_state(path)
    assert isinstance(v2.style, R.IsDynamicSemistochastic) and np.array_equal(v2.vals, [0.5, 1.5, -2.0])
    v3, _ = R.load_state(path, style=R.IsDeterministic())
    assert isinstance(v3.style, R.IsDeterministic)
    # several ranks (_save_state_mpi, RimuIO.jl:107-135): rank 0 writes the streaming format with the metadata, the others
    # append their record batch in rank order between barriers; the file reads back as the concatenation in that order
    class RankCtx:
        def __init__(self, rank, nranks, log):
            self.rank, self.nranks, self.log = rank, nranks, log

        def allreduce(self, values):
            self.log.append(("barrier", self.rank))
            return values

    path2, log = tmp_path / "ranks.arrow", []
    shares = [(keys[[0]], np.array([1.0])), (keys[[1]], np.array([2.0])), (keys[[2]], np.array([3.0]))]
    # the ranks run one after the other here, which is the order the barriers enforce between real processes
    for r, (k, x) in enumerate(shares):
        R.save_state(path2, FakeVec(k, x, address_type=addrs[0].address_type, ctx=RankCtx(r, 3, log)), **({"step": 7} if r == 0 else {}))
    k_all, v_all, at, meta = rimuio.read_state_file(path2)
    assert np.array_equal(k_all, keys) and np.array_equal(v_all, [1.0, 2.0, 3.0]) and meta["step"] == 7
    assert sum(1 for e in log if e == ("barrier", 1)) == 3  # one barrier before every appending rank + the final one
    monkeypatch.setattr(FakeVec, "upload", lambda self, k, x: self.assign(k, x), raising=False)
    v4, _ = R.load_state(path2, ctx=RankCtx(0, 3, log))  # a multi-rank load goes through upload(), which keeps the owned keys only
    assert np.array_equal(v4.vals, [1.0, 2.0, 3.0])


def test_general_composite_state_files(built, tmp_path):
    """CompositeFS with bosonic / mixed components: struct of the components' chunk lists (arrowtypes.jl:113-161), one and
    two device words; the extension metadata names every component's storage."""
    import rimu_b200 as R
    from rimu_b200 import rimuio
    from tests.cases import oracle_ham, product_ham, sample_keys
    for name, meta in (("rs_comp_bf", b"Rimu.BoseFS.BitString:3.6.8;Rimu.FermiFS.BitString:1.6.6"),
                       ("rs_comp_ffb", b"Rimu.FermiFS.BitString:2.4.4;Rimu.FermiFS.BitString:2.4.4;Rimu.BoseFS.BitString:3.4.6"),
                       ("rs_comp_bf_w2", b"Rimu.BoseFS.BitString:30.27.56;Rimu.FermiFS.BitString:5.27.27")):
        oh, ph = oracle_ham(name), product_ham(name)
        keys = sample_keys(oh, 12, seed=3)
        vals = np.arange(1, len(keys) + 1, dtype=np.float64) * 0.5
        path = tmp_path / (name + ".arrow")
        rimuio.write_state_file(path, keys, vals, ph.address_type)
        tbl = _read_raw(path)
        kf = tbl.schema.field("key")
        assert kf.metadata[b"ARROW:extension:name"] == b"Rimu.CompositeFS" and kf.metadata[b"ARROW:extension:metadata"] == meta
        col = tbl.column("key").combine_chunks()
        # component 1 of the first key, as the reference stores it: the component's own bit string
        comp0 = ph.address_type.from_key(keys[0]).components[0]
        assert int(col.field(0).flatten().to_numpy()[0]) == comp0._bits()
        k2, v2, at2, _ = rimuio.read_state_file(path)
        assert at2 == ph.address_type and np.array_equal(k2, keys) and np.array_equal(v2, vals)
