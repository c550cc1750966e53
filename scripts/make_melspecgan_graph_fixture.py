"""Extracts the facts the reference's exported MelspecGAN inference graph pins (models/melspecgan/infer.meta,
a serialised MetaGraphDef written by models/melspecgan/infer.py) into tests/golden/melspecgan_infer_graph.json:

  * every variable (VariableV2 node): name, dtype enum, shape  -> checkpoint names and filter layouts
  * every Conv2DBackpropInput / FusedBatchNorm / MatMul / Reshape / activation node of the G scope: op,
    inputs and the attributes that fix the arithmetic (strides, padding, data_format, epsilon, is_training)

Run in the authoring container (reads /root/reference); the JSON is committed, the GPU box never sees
the reference.  No TensorFlow / protobuf package needed: the wire format is decoded with the minimal
reader of advoc_b200.tf_bundle.
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from advoc_b200.tf_bundle import _proto_fields, _parse_shape, _varint  # noqa: E402

SRC = '/root/reference/models/melspecgan/infer.meta'
DST = os.path.join(ROOT, 'tests', 'golden', 'melspecgan_infer_graph.json')


def tensor_value(buf):
  """TensorProto (tensor.proto): dtype=1, tensor_shape=2, tensor_content=4, float_val=5, int_val=7."""
  dtype, shape, content, floats, ints = 0, (), b'', [], []
  for field, wt, val in _proto_fields(buf):
    if field == 1:
      dtype = val
    elif field == 2:
      shape = _parse_shape(val)
    elif field == 4:
      content = val
    elif field == 5:
      floats += list(struct.unpack('<%df' % (len(val) // 4), val)) if wt == 2 else [struct.unpack('<f', struct.pack('<I', val))[0]]
    elif field == 7:
      if wt == 2:
        pos = 0
        while pos < len(val):
          x, pos = _varint(val, pos)
          ints.append(x - (1 << 64) if x >= (1 << 63) else x)
      else:
        ints.append(val - (1 << 64) if val >= (1 << 63) else val)
  if content:
    fmt = {1: 'f', 3: 'i', 9: 'q'}.get(dtype)
    if fmt:
      n = len(content) // struct.calcsize(fmt)
      return list(struct.unpack('<%d%s' % (n, fmt), content))
    return None
  return floats or ints


def attr_value(buf):
  """AttrValue (attr_value.proto): s=2 bytes, i=3, f=4, b=5, type=6, shape=7, list=1 {s=2, i=3}."""
  out = {}
  for field, wt, val in _proto_fields(buf):
    if field == 2:
      out['s'] = val.decode('utf-8', 'replace')
    elif field == 3:
      out['i'] = val - (1 << 64) if val >= (1 << 63) else val
    elif field == 4:
      out['f'] = struct.unpack('<f', struct.pack('<I', val))[0]
    elif field == 5:
      out['b'] = bool(val)
    elif field == 6:
      out['type'] = val
    elif field == 7:
      out['shape'] = list(_parse_shape(val))
    elif field == 8:
      out['tensor'] = tensor_value(val)
    elif field == 1:
      ints, strs = [], []
      for f2, wt2, v2 in _proto_fields(val):
        if f2 == 3:
          if wt2 == 2:     # packed
            pos = 0
            while pos < len(v2):
              x, pos = _varint(v2, pos)
              ints.append(x)
          else:
            ints.append(v2)
        elif f2 == 2:
          strs.append(v2.decode('utf-8', 'replace'))
      out['list'] = ints if ints else strs
  return out


def main():
  meta = open(SRC, 'rb').read()
  graph = None
  for field, _, val in _proto_fields(meta):
    if field == 2:               # MetaGraphDef.graph_def
      graph = val
  nodes = []
  for field, _, val in _proto_fields(graph):
    if field != 1:               # GraphDef.node
      continue
    node = dict(name='', op='', inputs=[], attr={})
    for f2, _, v2 in _proto_fields(val):
      if f2 == 1:
        node['name'] = v2.decode()
      elif f2 == 2:
        node['op'] = v2.decode()
      elif f2 == 3:
        node['inputs'].append(v2.decode())
      elif f2 == 5:              # map<string, AttrValue> entry {key=1, value=2}
        k, a = None, {}
        for f3, _, v3 in _proto_fields(v2):
          if f3 == 1:
            k = v3.decode()
          elif f3 == 2:
            a = attr_value(v3)
        node['attr'][k] = a
    nodes.append(node)
  variables = [dict(name=n['name'], dtype=n['attr']['dtype']['type'], shape=n['attr']['shape']['shape'])
               for n in nodes if n['op'] in ('VariableV2', 'VarHandleOp')]
  keep_ops = ('Conv2DBackpropInput', 'Conv2D', 'FusedBatchNorm', 'FusedBatchNormV2', 'MatMul', 'BiasAdd', 'Relu',
              'Tanh', 'Reshape', 'Placeholder', 'RandomStandardNormal', 'Identity')
  keep_attr = ('strides', 'padding', 'data_format', 'epsilon', 'is_training', 'transpose_a', 'transpose_b',
               'dilations')
  ops = []
  for n in nodes:
    if n['op'] in keep_ops and (n['name'].startswith('G/') or n['op'] in ('Placeholder', 'RandomStandardNormal')
                                or n['name'] == 'G_z') and '/Initializer/' not in n['name'] \
        and not n['name'].endswith('/read') and '/Assign' not in n['name']:
      ops.append(dict(name=n['name'], op=n['op'], inputs=n['inputs'],
                      attr={k: (v.get('s', v.get('list', v.get('f', v.get('b', v.get('i'))))))
                            for k, v in n['attr'].items() if k in keep_attr}))
  consts = {}
  for n in nodes:            # reshape target, transposed-conv output sizes, feats_denorm constants
    if n['op'] == 'Const' and 'value' in n['attr'] and '/Initializer/' not in n['name'] and \
        (n['name'].startswith('G/') or n['name'] in ('add/y', 'mul/y', 'samp_z/shape/1')):
      consts[n['name']] = n['attr']['value'].get('tensor')
  tail = [dict(name=n['name'], op=n['op'], inputs=n['inputs']) for n in nodes if n['name'] in ('add', 'mul', 'G_z')]
  with open(DST, 'w') as f:
    json.dump(dict(source='models/melspecgan/infer.meta', n_nodes=len(nodes), variables=variables, ops=ops,
                   consts=consts, tail=tail), f, indent=1, sort_keys=True)
  print('wrote', DST, len(variables), 'variables', len(ops), 'ops of', len(nodes), 'nodes')


if __name__ == '__main__':
  main()
