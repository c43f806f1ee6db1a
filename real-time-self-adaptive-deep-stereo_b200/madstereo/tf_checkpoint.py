"""Stand-alone reader / writer for TensorFlow "V2" checkpoints (tensor bundles): `<prefix>.index` + `<prefix>.data-NNNNN-of-MMMMM`.

Replaces, for weight import, `tf.train.NewCheckpointReader` / `tf.train.Saver.restore` as used by the reference
(Data_utils/weights_utils.py:27-37, Stereo_Online_Adaptation.py:150-154, 243) -- TensorFlow is not a dependency of this
repository.  Host-side only (numpy); the arrays go to the engine through `OnlineAdaptation.load_weights`.

Format (tensorflow/core/util/tensor_bundle + lib/io/table, a fork of the LevelDB table format):
  * `.index` is an SSTable.  Footer = last 48 bytes: metaindex BlockHandle, index BlockHandle (varint64 offset, size
    each), zero padding to 40 bytes, 8-byte magic 0xdb4775248b80fb57 (little endian).  A block is a run of entries
    `varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value`, then a uint32 restart array and
    the uint32 restart count; it is followed on disk by a 1-byte compression type and a 4-byte masked CRC32C.  The
    index block maps separator keys to data-block handles.  Bundles are written uncompressed (type 0); snappy blocks
    (type 1) are rejected with a clear error.
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness (0 = little), 3: version};
    key <tensor name> -> BundleEntryProto {1: dtype, 2: TensorShapeProto {2: dim {1: size}}, 3: shard_id, 4: offset,
    5: size, 6: crc32c (fixed32, masked), 7: slices (partitioned variables: not supported)}.
  * tensor bytes are raw little-endian row-major values at [offset, offset+size) of the shard file.

Validated by round trips through the writer below and by the checks the format itself offers (block and tensor CRC32C);
no TensorFlow-written file is available in the build environment, so compatibility with real checkpoints rests on the
format description above.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(Exception):
    pass


# ------------------------------------------------------------------------------------------------
# CRC32C (Castagnoli), masked the LevelDB way
# ------------------------------------------------------------------------------------------------
def _crc_table():
    poly = 0x82F63B78
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t[i] = c
    return t


_TABLE = _crc_table()


def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def _shift_operator(nbytes):
    """32x32 GF(2) matrix that advances a (reflected, Castagnoli) CRC register over `nbytes` zero bytes -- zlib's
    crc32_combine construction."""
    op = [0x82F63B78] + [1 << n for n in range(31)]          # one zero bit
    op = _gf2_square(_gf2_square(_gf2_square(op)))           # 8 zero bits = one byte
    result = None
    n = nbytes
    while n:
        if n & 1:
            result = op if result is None else [_gf2_times(op, result[k]) for k in range(32)]
        n >>= 1
        if n:
            op = _gf2_square(op)
    return result if result is not None else [1 << k for k in range(32)]


def _crc_scalar(buf, c):
    tab = _TABLE
    for b in buf.tolist():
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c


def crc32c(data, crc=0):
    """CRC32C of bytes / a uint8 array.  Long inputs are cut into 4096 equal lanes whose table-driven CRCs advance together
    as numpy vectors (one iteration per byte POSITION, not per byte) and are then chained with the zero-shift operator
    (crc(A||B) = shift_len(B)(crc(A)) xor crc(B)): a 2 MB tensor takes milliseconds instead of a second."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).ravel()
    n = buf.size
    lanes = 4096
    if n < 64 * lanes:
        return _crc_scalar(buf, crc ^ 0xFFFFFFFF) ^ 0xFFFFFFFF
    seg = n // lanes
    body = buf[:seg * lanes].reshape(lanes, seg)
    c = np.full(lanes, 0xFFFFFFFF, dtype=np.uint32)
    c[0] = np.uint32(crc ^ 0xFFFFFFFF)
    tab = _TABLE
    for i in range(seg):
        c = tab[(c ^ body[:, i]) & np.uint32(0xFF)] ^ (c >> np.uint32(8))
    finals = (c ^ np.uint32(0xFFFFFFFF)).tolist()             # lane 0 carries the incoming crc, the others start fresh
    op = _shift_operator(seg)
    total = finals[0]
    for f in finals[1:]:
        total = _gf2_times(op, total) ^ f
    return _crc_scalar(buf[seg * lanes:], total ^ 0xFFFFFFFF) ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - 0xa282ead8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------
# varints / tiny protobuf codec
# ------------------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError('truncated varint')
        b = buf[pos]; pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError('varint too long')


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """{field number: [values]}; values are ints (varint / fixed) or bytes (length-delimited)."""
    fields, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n]); pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]; pos += 4
        else:
            raise CheckpointError('unsupported protobuf wire type %d' % wt)
        fields.setdefault(num, []).append(v)
    return fields


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


# ------------------------------------------------------------------------------------------------
# table reader
# ------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    if offset + size + 5 > len(data):
        raise CheckpointError('block handle outside the index file')
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from('<I', data, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
            raise CheckpointError('index block checksum mismatch')
    if ctype != 0:
        raise CheckpointError('compressed index blocks (type %d) are not supported' % ctype)
    return raw


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError('malformed block')
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared]); pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


class CheckpointReader:
    """Subset of tf.train.NewCheckpointReader: has_tensor / get_variable_to_shape_map / get_variable_to_dtype_map /
    get_tensor.  `prefix` is what the reference passes as --weights (the path without .index / .data-...)."""

    def __init__(self, prefix, verify=False):
        self.prefix = prefix
        self.verify = verify
        index = prefix + '.index'
        if not os.path.exists(index):
            raise CheckpointError('checkpoint index %s not found' % index)
        data = open(index, 'rb').read()
        if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != _MAGIC:
            raise CheckpointError('%s is not a TensorFlow V2 checkpoint index (bad magic)' % index)
        footer = data[len(data) - 48:]
        _mo, p = _get_varint(footer, 0); _ms, p = _get_varint(footer, p)
        io, p = _get_varint(footer, p); isz, p = _get_varint(footer, p)
        self._entries = {}
        self.header = None
        for _key, handle in _block_entries(_read_block(data, io, isz, verify)):
            bo, q = _get_varint(handle, 0); bs, q = _get_varint(handle, q)
            for key, value in _block_entries(_read_block(data, bo, bs, verify)):
                if key == b'':
                    h = _parse_proto(value)
                    self.header = {'num_shards': h.get(1, [1])[0], 'endianness': h.get(2, [0])[0]}
                else:
                    self._entries[key.decode('utf-8')] = self._entry(value)
        if self.header is None:
            raise CheckpointError('bundle header missing')
        if self.header['endianness'] != 0:
            raise CheckpointError('big-endian bundles are not supported')

    @staticmethod
    def _entry(value):
        f = _parse_proto(value)
        shape = []
        if 2 in f:
            for dim in _parse_proto(f[2][0]).get(2, []):
                shape.append(_signed64(_parse_proto(dim).get(1, [0])[0]))
        return {'dtype': f.get(1, [0])[0], 'shape': tuple(shape), 'shard': f.get(3, [0])[0], 'offset': f.get(4, [0])[0],
                'size': f.get(5, [0])[0], 'crc': f.get(6, [None])[0], 'sliced': 7 in f}

    def has_tensor(self, name):
        return name in self._entries

    def get_variable_to_shape_map(self):
        return {k: list(v['shape']) for k, v in self._entries.items()}

    def get_variable_to_dtype_map(self):
        return {k: _DTYPES.get(v['dtype']) for k, v in self._entries.items()}

    def get_tensor(self, name):
        if name not in self._entries:
            raise CheckpointError('tensor %s not in checkpoint %s' % (name, self.prefix))
        e = self._entries[name]
        if e['sliced']:
            raise CheckpointError('%s is a partitioned variable (slices are not supported)' % name)
        if e['dtype'] not in _DTYPES:
            raise CheckpointError('%s: unsupported dtype enum %d' % (name, e['dtype']))
        shard = '%s.data-%05d-of-%05d' % (self.prefix, e['shard'], self.header['num_shards'])
        if not os.path.exists(shard):
            raise CheckpointError('checkpoint shard %s not found' % shard)
        with open(shard, 'rb') as f:
            f.seek(e['offset'])
            raw = f.read(e['size'])
        dt = np.dtype(_DTYPES[e['dtype']])
        count = int(np.prod(e['shape'])) if e['shape'] else 1
        if len(raw) != e['size'] or count * dt.itemsize != e['size']:
            raise CheckpointError('%s: size mismatch (entry %d bytes, shape %s)' % (name, e['size'], e['shape']))
        if self.verify and e['crc'] is not None and unmask_crc(e['crc']) != crc32c(raw):
            raise CheckpointError('%s: tensor checksum mismatch' % name)
        return np.frombuffer(raw, dtype=dt.newbyteorder('<')).reshape(e['shape']).astype(dt, copy=True)


# ------------------------------------------------------------------------------------------------
# writer (single shard, uncompressed) -- used for export and by the tests
# ------------------------------------------------------------------------------------------------
def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_checkpoint(prefix, tensors, block_size=4096):
    """tensors: {name: array}.  Writes <prefix>.index and <prefix>.data-00000-of-00001."""
    names = sorted(tensors, key=lambda s: s.encode('utf-8'))
    entries, offset = [], 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for n in names:
            a = np.asarray(tensors[n])
            a = a if a.flags['C_CONTIGUOUS'] else a.copy()            # (np.ascontiguousarray would turn 0-d into 1-d)
            if a.dtype not in _DTYPE_ENUM:
                raise CheckpointError('%s: dtype %s cannot be stored' % (n, a.dtype))
            raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()
            f.write(raw)
            shape = b''.join(_field(2, 2, _put_varint(len(d)) + d) for d in [_field(1, 0, _put_varint(int(s))) for s in a.shape])
            entry = (_field(1, 0, _put_varint(_DTYPE_ENUM[a.dtype])) + _field(2, 2, _put_varint(len(shape)) + shape) +
                     _field(4, 0, _put_varint(offset)) + _field(5, 0, _put_varint(len(raw))) +
                     _field(6, 5, struct.pack('<I', mask_crc(crc32c(raw)))))
            entries.append((n.encode('utf-8'), entry))
            offset += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))
    items = [(b'', header)] + entries
    out = bytearray()

    def emit(block):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block); out.append(0)
        out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        return handle

    index_items, cur, cur_bytes = [], [], 0
    for k, v in items:
        cur.append((k, v)); cur_bytes += len(k) + len(v) + 6
        if cur_bytes >= block_size:
            index_items.append((cur[-1][0], emit(_build_block(cur)))); cur, cur_bytes = [], 0
    if cur:
        index_items.append((cur[-1][0], emit(_build_block(cur))))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block(index_items, restart_interval=1))
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', _MAGIC)
    out.extend(footer)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
