"""Hand-assembles a TensorFlow V2 checkpoint (TensorBundle) byte by byte from the published formats -- LevelDB table
format (table_format.md: prefix-compressed entries, restart array, 5-byte block trailer = type + masked crc32c, 48-byte
footer with the magic 0xdb4775248b80fb57) and tensor_bundle.proto / tensor_shape.proto / versions.proto -- WITHOUT using
madstereo/tf_checkpoint.py (its writer must not validate its own reader).  Differences from what that writer emits, on
purpose: shared key prefixes with a restart interval of 16 inside ONE data block plus a second data block, a header with
an explicit endianness field, shard_id written explicitly, a scalar and an int32 tensor.

  python tests/golden/make_tensorbundle_fixture.py   -> tests/golden/tensorbundle_handmade.{index,data-00000-of-00001,json}
"""
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def crc32c(data):
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ 0x82F63B78 if crc & 1 else crc >> 1
    return crc ^ 0xFFFFFFFF


def masked(crc):
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def pb_varint(field, v):
    return varint(field << 3 | 0) + varint(v)


def pb_bytes(field, payload):
    return varint(field << 3 | 2) + varint(len(payload)) + payload


def pb_fixed32(field, v):
    return varint(field << 3 | 5) + struct.pack('<I', v)


def shape_proto(shape):              # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
    return b''.join(pb_bytes(2, pb_varint(1, int(s))) for s in shape)


def block(entries, restart_interval):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        last = k
    for r in restarts or [0]:
        out += struct.pack('<I', r)
    out += struct.pack('<I', max(1, len(restarts)))
    return bytes(out)


def main():
    rng = np.random.default_rng(20260923)
    tensors = {
        'model/gc-read-pyramid/conv1/biases': rng.standard_normal(16).astype('<f4'),
        'model/gc-read-pyramid/conv1/weights': rng.standard_normal((3, 3, 3, 16)).astype('<f4'),
        'model/gc-read-pyramid/conv2/biases': rng.standard_normal(16).astype('<f4'),
        'model/gc-read-pyramid/conv2/weights': rng.standard_normal((3, 3, 16, 16)).astype('<f4'),
        'global_step': np.array(4200, dtype='<i4'),
        'model/context-7/weights': rng.standard_normal((3, 3, 32, 1)).astype('<f4'),
    }
    DT = {np.dtype('<f4'): 1, np.dtype('<i4'): 3}     # types.proto: DT_FLOAT = 1, DT_INT32 = 3
    names = sorted(tensors, key=lambda s: s.encode())
    data, entries = bytearray(), []
    for n in names:
        a = tensors[n]
        raw = a.tobytes()
        entry = (pb_varint(1, DT[a.dtype]) + pb_bytes(2, shape_proto(a.shape)) + pb_varint(3, 0) + pb_varint(4, len(data)) +
                 pb_varint(5, len(raw)) + pb_fixed32(6, masked(crc32c(raw))))
        entries.append((n.encode(), entry))
        data += raw
    header = pb_varint(1, 1) + pb_varint(2, 0) + pb_bytes(3, pb_varint(1, 1))      # num_shards 1, LITTLE, VersionDef{producer 1}
    items = [(b'', header)] + entries

    index = bytearray()

    def emit(b):
        handle = varint(len(index)) + varint(len(b))
        index.extend(b)
        index.append(0)                                                     # kNoCompression
        index.extend(struct.pack('<I', masked(crc32c(b + b'\x00'))))
        return handle

    first, second = items[:4], items[4:]                                    # two data blocks
    h1 = emit(block(first, 16))
    h2 = emit(block(second, 16))
    meta = emit(block([], 16))
    idx = emit(block([(first[-1][0], h1), (second[-1][0], h2)], 1))
    footer = meta + idx
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    index.extend(footer)

    prefix = os.path.join(HERE, 'tensorbundle_handmade')
    open(prefix + '.index', 'wb').write(bytes(index))
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    json.dump({n: {'shape': list(a.shape), 'dtype': str(a.dtype), 'values': a.ravel().tolist()} for n, a in tensors.items()},
              open(prefix + '.json', 'w'))
    print('wrote', prefix, len(index), 'index bytes,', len(data), 'data bytes')


if __name__ == '__main__':
    main()
