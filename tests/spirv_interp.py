"""A small SPIR-V interpreter: just enough of the specification to EXECUTE the seven shader modules the
reference ships (`/root/reference/shaders/*.spv`, built from `shaders/*.glsl` by glslang and from
`shaders/ray-tracing/src/lib.rs` by rust-gpu).  TEST INFRASTRUCTURE: it lets the test-suite run the
reference's own compiled shading arithmetic on this CPU and pin `oracle/` (and through it the CUDA
path) against it — SURVEY.md 8c calls the shipped .spv files the only executable artefact of the
reference that can run here.

What is interpreted: every arithmetic, logic, conversion, composite, memory and control-flow
instruction the modules contain, one fp32 rounding per instruction (numpy float32), GLSL.std.450
extended instructions in fp32.  What is a callback (the parts that live in the Vulkan driver / RT
hardware, not in the shaders): `OpTraceRayKHR`, `OpImageSampleExplicitLod`, `OpImageWrite`,
`OpReadClockKHR`, loads through PhysicalStorageBuffer pointers (`Memory`), and the built-in inputs.

Written from the SPIR-V 1.4 specification's instruction layouts; nothing here comes from the reference.
"""
import math
import struct

import numpy as np

F32 = np.float32
U32 = np.uint32
I32 = np.int32
U64 = np.uint64

OPNAMES = {
    1: "Undef", 3: "Source", 4: "SourceExtension", 5: "Name", 6: "MemberName", 10: "Extension", 11: "ExtInstImport", 12: "ExtInst",
    14: "MemoryModel", 15: "EntryPoint", 16: "ExecutionMode", 17: "Capability", 19: "TypeVoid", 20: "TypeBool", 21: "TypeInt",
    22: "TypeFloat", 23: "TypeVector", 24: "TypeMatrix", 25: "TypeImage", 26: "TypeSampler", 27: "TypeSampledImage", 28: "TypeArray",
    29: "TypeRuntimeArray", 30: "TypeStruct", 32: "TypePointer", 33: "TypeFunction", 39: "TypeForwardPointer", 41: "ConstantTrue",
    42: "ConstantFalse", 43: "Constant", 44: "ConstantComposite", 46: "ConstantNull", 54: "Function", 55: "FunctionParameter",
    56: "FunctionEnd", 57: "FunctionCall", 59: "Variable", 61: "Load", 62: "Store", 63: "CopyMemory", 65: "AccessChain",
    66: "InBoundsAccessChain", 71: "Decorate", 72: "MemberDecorate", 79: "VectorShuffle", 80: "CompositeConstruct",
    81: "CompositeExtract", 82: "CompositeInsert", 83: "CopyObject", 84: "Transpose", 87: "ImageSampleImplicitLod",
    88: "ImageSampleExplicitLod", 99: "ImageWrite", 109: "ConvertFToU", 110: "ConvertFToS", 111: "ConvertSToF", 112: "ConvertUToF",
    113: "UConvert", 114: "SConvert", 120: "ConvertUToPtr", 124: "Bitcast", 126: "SNegate", 127: "FNegate", 128: "IAdd", 129: "FAdd",
    130: "ISub", 131: "FSub", 132: "IMul", 133: "FMul", 134: "UDiv", 135: "SDiv", 136: "FDiv", 137: "UMod", 138: "SRem", 139: "SMod",
    140: "FRem", 141: "FMod", 142: "VectorTimesScalar", 143: "MatrixTimesScalar", 144: "VectorTimesMatrix", 145: "MatrixTimesVector",
    146: "MatrixTimesMatrix", 148: "Dot", 164: "LogicalEqual", 165: "LogicalNotEqual", 166: "LogicalOr", 167: "LogicalAnd",
    168: "LogicalNot", 169: "Select", 170: "IEqual", 171: "INotEqual", 172: "UGreaterThan", 173: "SGreaterThan",
    174: "UGreaterThanEqual", 175: "SGreaterThanEqual", 176: "ULessThan", 177: "SLessThan", 178: "ULessThanEqual",
    179: "SLessThanEqual", 180: "FOrdEqual", 181: "FUnordEqual", 182: "FOrdNotEqual", 183: "FUnordNotEqual", 184: "FOrdLessThan",
    185: "FUnordLessThan", 186: "FOrdGreaterThan", 187: "FUnordGreaterThan", 188: "FOrdLessThanEqual", 189: "FUnordLessThanEqual",
    190: "FOrdGreaterThanEqual", 191: "FUnordGreaterThanEqual", 194: "ShiftRightLogical", 195: "ShiftRightArithmetic",
    196: "ShiftLeftLogical", 197: "BitwiseOr", 198: "BitwiseXor", 199: "BitwiseAnd", 200: "Not", 245: "Phi", 246: "LoopMerge",
    247: "SelectionMerge", 248: "Label", 249: "Branch", 250: "BranchConditional", 251: "Switch", 253: "Return", 254: "ReturnValue",
    255: "Unreachable", 400: "CopyLogical", 4445: "TraceRayKHR", 4447: "ConvertUToAccelerationStructureKHR",
    4448: "IgnoreIntersectionKHR", 4449: "TerminateRayKHR", 5056: "ReadClockKHR", 5341: "TypeAccelerationStructureKHR",
}

# decorations / built-ins / storage classes used below
DEC_BUILTIN, DEC_ARRAY_STRIDE, DEC_MATRIX_STRIDE, DEC_OFFSET, DEC_COL_MAJOR, DEC_ROW_MAJOR = 11, 6, 7, 35, 5, 4
DEC_BINDING, DEC_DESCRIPTOR_SET = 33, 34
BUILTINS = {
    5319: "LaunchId", 5320: "LaunchSize", 5321: "WorldRayOrigin", 5322: "WorldRayDirection", 5323: "ObjectRayOrigin",
    5324: "ObjectRayDirection", 5325: "RayTmin", 5326: "RayTmax", 5327: "InstanceCustomIndex", 5330: "ObjectToWorld",
    5331: "WorldToObject", 5332: "HitT", 5333: "HitKind", 5351: "IncomingRayFlags", 5352: "RayGeometryIndex", 6: "InstanceId",
    7: "PrimitiveId",
}
SC_UNIFORM_CONSTANT, SC_INPUT, SC_UNIFORM, SC_PRIVATE, SC_FUNCTION, SC_PUSH_CONSTANT, SC_STORAGE_BUFFER = 0, 1, 2, 6, 7, 9, 12
SC_RAY_PAYLOAD, SC_HIT_ATTRIBUTE, SC_INCOMING_RAY_PAYLOAD, SC_PHYSICAL = 5338, 5339, 5342, 5349


class IgnoreIntersection(Exception):
    """OpIgnoreIntersectionKHR terminated the any-hit invocation."""


class SpirvError(Exception):
    pass


class Type:
    __slots__ = ("kind", "width", "signed", "elem", "count", "members", "storage", "id")

    def __init__(self, kind, **kw):
        self.kind = kind
        self.width = self.signed = self.elem = self.count = self.members = self.storage = self.id = None
        for k, v in kw.items():
            setattr(self, k, v)


class Memory:
    """Byte-addressed device memory behind PhysicalStorageBuffer pointers: regions registered at 64-bit addresses."""

    def __init__(self):
        self.regions = []  # (base, bytes-like numpy uint8)
        self.next = 0x10000000

    def add(self, array, base=None):
        data = np.ascontiguousarray(array).view(np.uint8).reshape(-1)
        if base is None:
            base = self.next
            self.next = (base + len(data) + 0xFFF + 0x1000) & ~0xFFF
        self.regions.append((base, data))
        return base

    def read(self, addr, fmt, size):
        addr = int(addr)
        for base, data in self.regions:
            if base <= addr and addr + size <= base + len(data):
                return struct.unpack_from(fmt, data, addr - base)[0]
        raise SpirvError(f"load of {size} bytes at unmapped address {addr:#x}")


class PhysPtr:
    """PhysicalStorageBuffer pointer; `layout` = the decorations of the struct member it points into (matrix strides)."""
    __slots__ = ("addr", "type", "layout")

    def __init__(self, addr, type_, layout=None):
        self.addr, self.type, self.layout = int(addr), type_, layout


class VarPtr:
    """Logical pointer: a cell (python list of length 1 holding the root value) plus an index path."""
    __slots__ = ("cell", "path", "type")

    def __init__(self, cell, path, type_):
        self.cell, self.path, self.type = cell, path, type_


class Module:
    def __init__(self, words):
        if isinstance(words, (bytes, bytearray)):
            words = np.frombuffer(words, dtype="<u4")
        self.words = [int(w) for w in words]
        if self.words[0] != 0x07230203:
            raise SpirvError("not a SPIR-V module")
        self.types, self.consts, self.names, self.member_names = {}, {}, {}, {}
        self.decor, self.member_decor = {}, {}
        self.globals = {}     # id -> (pointer type, storage class)
        self.ext_sets = {}
        self.entry = None     # (model, function id, name, interface ids)
        self.functions = {}   # id -> {"params": [...], "blocks": {label: [insts]}, "order": [labels], "first": label}
        self.undefs = {}
        self._parse()

    # ------------------------------------------------------------------ parsing
    @staticmethod
    def _string(ws):
        b = b"".join(struct.pack("<I", w) for w in ws)
        return b.split(b"\0", 1)[0].decode(), len(b.split(b"\0", 1)[0]) // 4 + 1

    def instructions(self):
        i, w = 5, self.words
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            if wc == 0:
                raise SpirvError("zero word count")
            yield op, w[i + 1:i + wc]
            i += wc

    def _parse(self):
        T = self.types
        cur, block = None, None
        for op, a in self.instructions():
            if op == 5:
                self.names[a[0]] = self._string(a[1:])[0]
            elif op == 6:
                self.member_names[(a[0], a[1])] = self._string(a[2:])[0]
            elif op == 11:
                self.ext_sets[a[0]] = self._string(a[1:])[0]
            elif op == 15:
                name, n = self._string(a[2:])
                self.entry = (a[0], a[1], name, list(a[2 + n:]))
            elif op == 71:
                self.decor.setdefault(a[0], {})[a[1]] = list(a[2:])
            elif op == 72:
                self.member_decor.setdefault((a[0], a[1]), {})[a[2]] = list(a[3:])
            elif op == 19:
                T[a[0]] = Type("void", id=a[0])
            elif op == 20:
                T[a[0]] = Type("bool", id=a[0])
            elif op == 21:
                T[a[0]] = Type("int", width=a[1], signed=bool(a[2]), id=a[0])
            elif op == 22:
                T[a[0]] = Type("float", width=a[1], id=a[0])
            elif op == 23:
                T[a[0]] = Type("vector", elem=T[a[1]], count=a[2], id=a[0])
            elif op == 24:
                T[a[0]] = Type("matrix", elem=T[a[1]], count=a[2], id=a[0])  # `count` columns of vector `elem`
            elif op == 25:
                T[a[0]] = Type("image", id=a[0])
            elif op == 26:
                T[a[0]] = Type("sampler", id=a[0])
            elif op == 27:
                T[a[0]] = Type("sampled_image", id=a[0])
            elif op == 28:
                T[a[0]] = Type("array", elem=T[a[1]], count=("const", a[2]), id=a[0])
            elif op == 29:
                T[a[0]] = Type("array", elem=T[a[1]], count=None, id=a[0])
            elif op == 30:
                T[a[0]] = Type("struct", members=[T[m] for m in a[1:]], id=a[0])
            elif op == 39:
                T[a[0]] = Type("pointer", storage=a[1], elem=None, id=a[0])
            elif op == 32:
                if a[0] in T:  # completes a forward pointer
                    T[a[0]].elem = T[a[2]]
                else:
                    T[a[0]] = Type("pointer", storage=a[1], elem=T[a[2]], id=a[0])
            elif op == 33:
                T[a[0]] = Type("function", id=a[0])
            elif op == 5341:
                T[a[0]] = Type("accel", id=a[0])
            elif op == 41:
                self.consts[a[1]] = True
            elif op == 42:
                self.consts[a[1]] = False
            elif op == 43:
                self.consts[a[1]] = self._scalar_const(T[a[0]], a[2:])
            elif op == 44:
                self.consts[a[1]] = self._compose(T[a[0]], [self.consts[c] for c in a[2:]])
            elif op == 46:
                self.consts[a[1]] = self.zero(T[a[0]])
            elif op == 1 and cur is None:
                self.consts[a[1]] = self.zero(T[a[0]])
            elif op == 59 and cur is None:
                self.globals[a[1]] = (T[a[0]], a[2], a[3] if len(a) > 3 else None)
            elif op == 54:
                cur = {"params": [], "blocks": {}, "order": [], "first": None, "result_type": T[a[0]]}
                self.functions[a[1]] = cur
            elif op == 55:
                cur["params"].append(a[1])
            elif op == 56:
                cur, block = None, None
            elif cur is not None:
                if op == 248:
                    block = []
                    cur["blocks"][a[0]] = block
                    cur["order"].append(a[0])
                    if cur["first"] is None:
                        cur["first"] = a[0]
                else:
                    block.append((op, tuple(a)))
        # resolve array lengths
        for t in T.values():
            if t.kind == "array" and isinstance(t.count, tuple):
                t.count = int(self.consts[t.count[1]])

    @staticmethod
    def _scalar_const(t, ws):
        if t.kind == "float":
            if t.width == 32:
                return F32(struct.unpack("<f", struct.pack("<I", ws[0]))[0])
            raise SpirvError("only 32-bit floats")
        if t.kind == "int":
            if t.width == 64:
                v = ws[0] | (ws[1] << 32)
                return np.int64(v - (1 << 64) if (t.signed and v >> 63) else v) if t.signed else U64(v)
            if t.width == 32:
                return I32(ws[0] - (1 << 32) if ws[0] >> 31 else ws[0]) if t.signed else U32(ws[0])
            if t.width == 8:
                return np.uint8(ws[0] & 0xFF) if not t.signed else np.int8(((ws[0] & 0xFF) ^ 0x80) - 0x80)
            if t.width == 16:
                return np.uint16(ws[0] & 0xFFFF) if not t.signed else np.int16(((ws[0] & 0xFFFF) ^ 0x8000) - 0x8000)
        raise SpirvError(f"constant of type {t.kind}")

    # ------------------------------------------------------------------ values
    @staticmethod
    def np_type(t):
        if t.kind == "float":
            return F32
        if t.kind == "bool":
            return np.bool_
        if t.kind == "int":
            return {(8, False): np.uint8, (8, True): np.int8, (16, False): np.uint16, (16, True): np.int16, (32, False): U32,
                    (32, True): I32, (64, False): U64, (64, True): np.int64}[(t.width, t.signed)]
        raise SpirvError(f"no numpy type for {t.kind}")

    def zero(self, t):
        if t.kind in ("float", "int"):
            return self.np_type(t)(0)
        if t.kind == "bool":
            return False
        if t.kind == "vector":
            return np.zeros(t.count, self.np_type(t.elem))
        if t.kind == "matrix":
            return [np.zeros(t.elem.count, F32) for _ in range(t.count)]
        if t.kind == "array":
            return [self.zero(t.elem) for _ in range(t.count or 0)]
        if t.kind == "struct":
            return [self.zero(m) for m in t.members]
        if t.kind == "pointer":
            return None
        return None

    def _compose(self, t, parts):
        if t.kind == "vector":
            flat = []
            for p in parts:
                if isinstance(p, np.ndarray):
                    flat.extend(p.tolist())
                else:
                    flat.append(p)
            return np.array(flat, self.np_type(t.elem))
        return [self.copy(p) for p in parts]

    @staticmethod
    def copy(v):
        if isinstance(v, np.ndarray):
            return v.copy()
        if isinstance(v, list):
            return [Module.copy(x) for x in v]
        return v

    # ------------------------------------------------------------------ layout of PhysicalStorageBuffer / Uniform data
    def size_of(self, t):
        if t.kind in ("float", "int"):
            return t.width // 8
        if t.kind == "vector":
            return t.count * self.size_of(t.elem)
        if t.kind == "pointer":
            return 8
        raise SpirvError(f"size_of({t.kind})")

    def load_phys(self, mem, addr, t, layout=None):
        k = t.kind
        if k == "float":
            return F32(mem.read(addr, "<f", 4))
        if k == "int":
            fmt = {(8, False): "<B", (8, True): "<b", (16, False): "<H", (16, True): "<h", (32, False): "<I", (32, True): "<i",
                   (64, False): "<Q", (64, True): "<q"}[(t.width, t.signed)]
            return self.np_type(t)(mem.read(addr, fmt, t.width // 8))
        if k == "bool":
            raise SpirvError("bool in a buffer")
        if k == "vector":
            s = self.size_of(t.elem)
            return np.array([self.load_phys(mem, addr + i * s, t.elem) for i in range(t.count)], self.np_type(t.elem))
        if k == "pointer":
            return PhysPtr(mem.read(addr, "<Q", 8), t.elem)
        if k == "struct":
            out = []
            for i, m in enumerate(t.members):
                md = self.member_decor.get((t.id, i), {})
                out.append(self.load_phys(mem, addr + md[DEC_OFFSET][0], m, md))
            return out
        if k == "array":
            stride = self.decor[t.id][DEC_ARRAY_STRIDE][0]
            if t.count is None:
                raise SpirvError("load of a whole runtime array")
            return [self.load_phys(mem, addr + i * stride, t.elem) for i in range(t.count)]
        if k == "matrix":
            stride = layout[DEC_MATRIX_STRIDE][0]
            if DEC_ROW_MAJOR in layout:
                rows = [[self.load_phys(mem, addr + r * stride + c * 4, t.elem.elem) for c in range(t.count)] for r in range(t.elem.count)]
                return [np.array([rows[r][c] for r in range(t.elem.count)], F32) for c in range(t.count)]
            return [self.load_phys(mem, addr + c * stride, t.elem) for c in range(t.count)]
        raise SpirvError(f"load_phys({k})")

    def phys_access(self, ptr, indices):
        addr, t, layout = ptr.addr, ptr.type, None
        for ix in indices:
            ix = int(ix)
            if t.kind == "struct":
                layout = self.member_decor.get((t.id, ix), {})
                addr += layout[DEC_OFFSET][0]
                t = t.members[ix]
            elif t.kind == "array":
                addr += ix * self.decor[t.id][DEC_ARRAY_STRIDE][0]
                t = t.elem
            elif t.kind == "matrix":
                if layout and DEC_ROW_MAJOR in layout:
                    raise SpirvError("access chain into a row-major matrix")
                addr += ix * layout[DEC_MATRIX_STRIDE][0]
                t = t.elem
            elif t.kind == "vector":
                addr += ix * self.size_of(t.elem)
                t = t.elem
            else:
                raise SpirvError(f"access chain into {t.kind}")
        return PhysPtr(addr, t, layout)

    def disassemble(self):
        out = []
        for op, a in self.instructions():
            out.append(f"{OPNAMES.get(op, op)} " + " ".join(str(x) for x in a))
        return "\n".join(out)


# ---------------------------------------------------------------------------------------------- GLSL.std.450
def _f(x):
    return np.asarray(x, F32)


def _normalize(v):
    v = _f(v)
    if v.ndim == 0:
        return F32(np.sign(v))
    n = F32(np.sqrt(F32(np.sum(v * v, dtype=F32))))
    return (v / n).astype(F32)


def _pow(x, y):
    with np.errstate(all="ignore"):
        return np.power(_f(x).astype(np.float64), _f(y).astype(np.float64)).astype(F32)


def _smoothstep(e0, e1, x):
    e0, e1, x = _f(e0), _f(e1), _f(x)
    t = np.clip(((x - e0) / (e1 - e0)).astype(F32), F32(0), F32(1)).astype(F32)
    return (t * t * (F32(3) - F32(2) * t)).astype(F32)


def _wrap(fn):
    def g(*a):
        with np.errstate(all="ignore"):
            r = fn(*a)
        return r
    return g


GLSL = {
    1: lambda x: np.round(_f(x)), 3: lambda x: np.trunc(_f(x)), 4: lambda x: np.abs(_f(x)), 6: lambda x: np.sign(_f(x)),
    8: lambda x: np.floor(_f(x)), 9: lambda x: np.ceil(_f(x)), 10: lambda x: (_f(x) - np.floor(_f(x))).astype(F32),
    13: lambda x: np.sin(_f(x).astype(np.float64)).astype(F32), 14: lambda x: np.cos(_f(x).astype(np.float64)).astype(F32),
    15: lambda x: np.tan(_f(x).astype(np.float64)).astype(F32),
    26: _pow, 27: lambda x: np.exp(_f(x).astype(np.float64)).astype(F32), 28: lambda x: np.log(_f(x).astype(np.float64)).astype(F32),
    29: lambda x: np.exp2(_f(x).astype(np.float64)).astype(F32), 30: lambda x: np.log2(_f(x).astype(np.float64)).astype(F32),
    31: lambda x: np.sqrt(_f(x)), 32: lambda x: (F32(1) / np.sqrt(_f(x).astype(np.float64))).astype(F32),
    37: lambda a, b: np.minimum(_f(a), _f(b)), 40: lambda a, b: np.maximum(_f(a), _f(b)),
    38: lambda a, b: np.minimum(a, b), 41: lambda a, b: np.maximum(a, b), 39: lambda a, b: np.minimum(a, b), 42: lambda a, b: np.maximum(a, b),
    43: lambda x, lo, hi: np.minimum(np.maximum(_f(x), _f(lo)), _f(hi)),
    44: lambda x, lo, hi: np.minimum(np.maximum(x, lo), hi), 45: lambda x, lo, hi: np.minimum(np.maximum(x, lo), hi),
    46: lambda x, y, a: (_f(x) * (F32(1) - _f(a)) + _f(y) * _f(a)).astype(F32),
    48: lambda edge, x: np.where(_f(x) < _f(edge), F32(0), F32(1)).astype(F32), 49: _smoothstep,
    50: lambda a, b, c: (_f(a).astype(np.float64) * _f(b).astype(np.float64) + _f(c).astype(np.float64)).astype(F32),
    66: lambda v: F32(np.sqrt(F32(np.sum(_f(v) * _f(v), dtype=F32)))) if _f(v).ndim else F32(abs(v)),
    67: lambda a, b: F32(np.sqrt(F32(np.sum((_f(a) - _f(b)) ** 2, dtype=F32)))),
    68: lambda a, b: np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F32),
    69: _normalize,
    71: lambda i, n: (_f(i) - F32(2) * F32(np.sum(_f(n) * _f(i), dtype=F32)) * _f(n)).astype(F32),
    79: lambda a, b: np.fmin(_f(a), _f(b)), 80: lambda a, b: np.fmax(_f(a), _f(b)),
    81: lambda x, lo, hi: np.fmin(np.fmax(_f(x), _f(lo)), _f(hi)),
}
GLSL = {k: _wrap(v) for k, v in GLSL.items()}


def _scalarize(v):
    """0-d numpy arrays -> numpy scalars (so that isinstance(v, np.ndarray) means 'vector')."""
    if isinstance(v, np.ndarray) and v.ndim == 0:
        return v[()]
    return v


# ---------------------------------------------------------------------------------------------- execution
class Invocation:
    """One shader invocation.  `env` supplies the callbacks:
         env.builtin(name) -> value                      built-in input variables
         env.trace_ray(accel, flags, cull, sbt_offset, sbt_stride, miss_index, origin, tmin, direction, tmax, payload_ptr)
         env.sample(image_index_or_handle, coord, lod) -> vec4
         env.image_write(image, coord, texel)
         env.read_clock() -> u64
       and the interface storage: env.memory (Memory), env.push_constants / env.uniform_buffers / env.payloads ...
       through `bind_global` below."""

    def __init__(self, module, env):
        self.m, self.env = module, env
        self.cells = {}  # global variable id -> cell
        self.steps = 0
        self.max_steps = 200000

    def bind(self, var_id, cell):
        self.cells[var_id] = cell

    def global_by_name(self, name):
        for vid in self.m.globals:
            if self.m.names.get(vid) == name:
                return vid
        raise KeyError(name)

    def globals_by_storage(self, sc):
        return [vid for vid, (t, s, _) in self.m.globals.items() if s == sc]

    # ---- pointers
    def _load(self, ptr):
        if isinstance(ptr, PhysPtr):
            return self.m.load_phys(self.env.memory, ptr.addr, ptr.type, ptr.layout)
        v = ptr.cell[0]
        for ix in ptr.path:
            v = v[ix]
        return Module.copy(v)

    def _store(self, ptr, value):
        if isinstance(ptr, PhysPtr):
            raise SpirvError("store through a physical pointer (the shaders only read buffers)")
        value = Module.copy(value)
        if not ptr.path:
            ptr.cell[0] = value
            return
        v = ptr.cell[0]
        for ix in ptr.path[:-1]:
            v = v[ix]
        v[ptr.path[-1]] = value

    def run(self, fn_id=None, args=()):
        m = self.m
        fn = m.functions[fn_id if fn_id is not None else m.entry[1]]
        V = dict(m.consts)
        for vid, (pt, sc, init) in m.globals.items():
            if vid not in self.cells:
                if sc == SC_INPUT:
                    b = m.decor.get(vid, {}).get(DEC_BUILTIN)
                    if b is None:
                        raise SpirvError(f"input variable {vid} without BuiltIn")
                    self.cells[vid] = [self.env.builtin(BUILTINS[b[0]], pt.elem)]
                elif sc in (SC_PRIVATE, SC_FUNCTION, SC_RAY_PAYLOAD, SC_HIT_ATTRIBUTE, SC_INCOMING_RAY_PAYLOAD):
                    self.cells[vid] = [m.copy(m.consts[init]) if init is not None else m.zero(pt.elem)]
                else:
                    cell = self.env.bind_global(m, vid, pt, sc)
                    self.cells[vid] = cell
            V[vid] = VarPtr(self.cells[vid], (), pt.elem)
        for p, a in zip(fn["params"], args):
            V[p] = a
        return self._exec(fn, V)

    def _exec(self, fn, V):
        m, env = self.m, self.env
        T = m.types
        label, prev = fn["first"], None
        blocks = fn["blocks"]
        with np.errstate(all="ignore"):
            while True:
                insts = blocks[label]
                # phis first (all read the values of the predecessor edge)
                phi_vals = []
                for op, a in insts:
                    if op != 245:
                        break
                    for k in range(2, len(a), 2):
                        if a[k + 1] == prev:
                            phi_vals.append((a[1], V[a[k]]))
                            break
                    else:
                        raise SpirvError("phi without matching predecessor")
                for rid, val in phi_vals:
                    V[rid] = val
                nxt = None
                for op, a in insts:
                    self.steps += 1
                    if self.steps > self.max_steps:
                        raise SpirvError(f"more than {self.max_steps} instructions in one invocation (runaway loop?)")
                    if op == 245 or op == 246 or op == 247:
                        continue
                    if op == 249:
                        nxt = a[0]
                        break
                    if op == 250:
                        nxt = a[1] if bool(V[a[0]]) else a[2]
                        break
                    if op == 251:
                        sel = int(V[a[0]])
                        nxt = a[1]
                        for k in range(2, len(a), 2):
                            if a[k] == sel & 0xFFFFFFFF:
                                nxt = a[k + 1]
                                break
                        break
                    if op == 253:
                        return None
                    if op == 254:
                        return V[a[0]]
                    if op == 255:
                        raise SpirvError("OpUnreachable executed")
                    if op == 4448:
                        raise IgnoreIntersection()
                    self._inst(op, a, V, T)
                if nxt is None:
                    raise SpirvError("block without terminator")
                prev, label = label, nxt

    # ---- one non-terminator instruction
    def _inst(self, op, a, V, T):
        m, env = self.m, self.env
        if op == 61:  # Load
            V[a[1]] = self._load(V[a[2]])
        elif op == 62:  # Store
            self._store(V[a[0]], V[a[1]])
        elif op == 63:  # CopyMemory
            self._store(V[a[0]], self._load(V[a[1]]))
        elif op == 59:  # Variable (Function storage)
            pt = T[a[0]]
            init = m.copy(V[a[3]]) if len(a) > 3 else m.zero(pt.elem)
            V[a[1]] = VarPtr([init], (), pt.elem)
        elif op in (65, 66):  # AccessChain
            base = V[a[2]]
            idx = [int(V[i]) for i in a[3:]]
            if isinstance(base, PhysPtr):
                V[a[1]] = m.phys_access(base, idx)
            else:
                t = base.type
                for ix in idx:
                    t = t.members[ix] if t.kind == "struct" else (t.elem if t.kind in ("array", "matrix") else t.elem)
                V[a[1]] = VarPtr(base.cell, base.path + tuple(idx), t)
        elif op == 120:  # ConvertUToPtr
            V[a[1]] = PhysPtr(int(V[a[2]]), T[a[0]].elem, None)
        elif op == 4447:  # ConvertUToAccelerationStructureKHR
            V[a[1]] = ("accel", int(V[a[2]]))
        elif op == 400 or op == 83:  # CopyLogical / CopyObject
            V[a[1]] = m.copy(V[a[2]])
        elif op == 1:  # Undef
            V[a[1]] = m.zero(T[a[0]])
        elif op == 12:  # ExtInst
            if m.ext_sets.get(a[2]) != "GLSL.std.450":
                raise SpirvError("unknown extended instruction set")
            fn = GLSL.get(a[3])
            if fn is None:
                raise SpirvError(f"GLSL.std.450 instruction {a[3]} not implemented")
            r = fn(*[V[x] for x in a[4:]])
            V[a[1]] = self._as(T[a[0]], r)
        elif op == 79:  # VectorShuffle
            v1, v2 = V[a[2]], V[a[3]]
            cat = np.concatenate([v1, v2])
            V[a[1]] = np.array([cat[c] if c != 0xFFFFFFFF else 0 for c in a[4:]], cat.dtype)
        elif op == 80:  # CompositeConstruct
            V[a[1]] = m._compose(T[a[0]], [V[x] for x in a[2:]])
        elif op == 81:  # CompositeExtract
            v = V[a[2]]
            for ix in a[3:]:
                v = v[ix]
            V[a[1]] = m.copy(v) if isinstance(v, (list, np.ndarray)) else v
        elif op == 82:  # CompositeInsert
            obj, comp = V[a[2]], m.copy(V[a[3]])
            v = comp
            for ix in a[4:-1]:
                v = v[ix]
            v[a[-1]] = obj
            V[a[1]] = comp
        elif op == 84:  # Transpose
            cols = V[a[2]]
            rows = len(cols[0])
            V[a[1]] = [np.array([cols[c][r] for c in range(len(cols))], F32) for r in range(rows)]
        elif op == 88:  # ImageSampleExplicitLod
            lod = V[a[5]] if len(a) > 5 and (a[4] & 2) else F32(0)
            V[a[1]] = np.asarray(env.sample(V[a[2]], V[a[3]], lod), F32)
        elif op == 99:  # ImageWrite
            env.image_write(V[a[0]], V[a[1]], V[a[2]])
        elif op == 4445:  # TraceRayKHR
            env.trace_ray(V[a[0]], int(V[a[1]]), int(V[a[2]]), int(V[a[3]]), int(V[a[4]]), int(V[a[5]]), V[a[6]], V[a[7]], V[a[8]], V[a[9]],
                          V[a[10]], self)
        elif op == 5056:  # ReadClockKHR
            V[a[1]] = self._as(T[a[0]], env.read_clock())
        else:
            V[a[1]] = self._as(T[a[0]], self._alu(op, a, V, T))

    def _as(self, t, r):
        """Coerce an ALU result to the declared result type (dtype + scalar/vector shape)."""
        if t.kind == "vector":
            if t.elem.kind == "bool":
                return np.asarray(r, np.bool_).reshape(t.count)
            return np.asarray(r).astype(Module.np_type(t.elem)).reshape(t.count)
        if t.kind == "bool":
            return bool(r)
        if t.kind in ("float", "int"):
            return Module.np_type(t)(np.asarray(r).reshape(-1)[0])
        return r

    def _alu(self, op, a, V, T):
        rt = T[a[0]]
        x = V[a[2]]
        y = V[a[3]] if len(a) > 3 else None
        if op == 129: return x + y
        if op == 131: return x - y
        if op == 133: return x * y
        if op == 136: return x / y
        if op == 127: return -x
        if op == 142: return (x * y).astype(F32)
        if op == 143: return [(c * y).astype(F32) for c in x]
        if op == 148: return F32(np.sum(x * y, dtype=F32))
        if op == 145:  # matrix * vector: sum of columns scaled
            acc = (x[0] * y[0]).astype(F32)
            for c in range(1, len(x)):
                acc = (acc + x[c] * y[c]).astype(F32)
            return acc
        if op == 144:  # vector * matrix
            return np.array([F32(np.sum(x * col, dtype=F32)) for col in y], F32)
        if op == 146:  # matrix * matrix
            out = []
            for col in y:
                acc = (x[0] * col[0]).astype(F32)
                for c in range(1, len(x)):
                    acc = (acc + x[c] * col[c]).astype(F32)
                out.append(acc)
            return out
        if op in (128, 130, 132):  # IAdd / ISub / IMul: modular arithmetic on python ints, any width
            st = rt.elem if rt.kind == "vector" else rt
            mask = (1 << st.width) - 1

            def wrap(p, q):
                r = (p + q if op == 128 else (p - q if op == 130 else p * q)) & mask
                return r - (1 << st.width) if (st.signed and r >> (st.width - 1)) else r
            if rt.kind == "vector":
                return [wrap(int(p), int(q)) for p, q in zip(x, y)]
            return wrap(int(x), int(y))
        if op == 126: return (-np.asarray(x).astype(np.int64))
        if op == 134 or op == 135: return np.asarray(x) // np.asarray(y) if op == 134 else np.trunc(np.asarray(x, np.float64) / np.asarray(y, np.float64))
        if op == 137: return np.asarray(x) % np.asarray(y)
        if op == 138: return np.fmod(np.asarray(x), np.asarray(y))
        if op == 139: return np.mod(np.asarray(x), np.asarray(y))
        if op == 140: return np.fmod(x, y)
        if op == 141: return (x - y * np.floor(x / y)).astype(F32)
        if op == 109 or op == 110: return np.trunc(np.asarray(x, np.float64))
        if op == 111 or op == 112: return np.asarray(x).astype(F32)
        if op == 113 or op == 114: return np.asarray(x)
        if op == 124:  # Bitcast
            if rt.kind == "pointer":
                return PhysPtr(int(x), rt.elem, None)
            if isinstance(x, PhysPtr):
                return U64(x.addr)
            src = np.asarray(x)
            dst = Module.np_type(rt.elem if rt.kind == "vector" else rt)
            return np.ascontiguousarray(src).reshape(-1).view(dst)
        if op == 164: return np.equal(x, y)
        if op == 165: return np.not_equal(x, y)
        if op == 166: return np.logical_or(x, y)
        if op == 167: return np.logical_and(x, y)
        if op == 168: return np.logical_not(x)
        if op == 169:
            z = V[a[4]]
            if isinstance(y, (list, PhysPtr, VarPtr)) or isinstance(z, (list, PhysPtr, VarPtr)):
                return y if bool(x) else z
            return np.where(x, y, z)
        if op == 170: return np.equal(x, y)
        if op == 171: return np.not_equal(x, y)
        if op in (172, 173): return np.greater(x, y)
        if op in (174, 175): return np.greater_equal(x, y)
        if op in (176, 177): return np.less(x, y)
        if op in (178, 179): return np.less_equal(x, y)
        if op == 180: return np.equal(x, y)
        if op == 182: return np.logical_and(np.not_equal(x, y), ~(np.isnan(x) | np.isnan(y)))
        if op == 183: return np.not_equal(x, y)
        if op == 181: return np.equal(x, y) | np.isnan(x) | np.isnan(y)
        if op == 184: return np.less(x, y)
        if op == 185: return ~np.greater_equal(x, y)
        if op == 186: return np.greater(x, y)
        if op == 187: return ~np.less_equal(x, y)
        if op == 188: return np.less_equal(x, y)
        if op == 189: return ~np.greater(x, y)
        if op == 190: return np.greater_equal(x, y)
        if op == 191: return ~np.less(x, y)
        if op == 194: return np.asarray(x) >> np.asarray(y).astype(np.asarray(x).dtype)
        if op == 195: return np.asarray(x) >> np.asarray(y).astype(np.asarray(x).dtype)
        if op == 196: return np.asarray(x) << np.asarray(y).astype(np.asarray(x).dtype)
        if op == 197: return np.asarray(x) | np.asarray(y).astype(np.asarray(x).dtype)
        if op == 198: return np.asarray(x) ^ np.asarray(y).astype(np.asarray(x).dtype)
        if op == 199: return np.asarray(x) & np.asarray(y).astype(np.asarray(x).dtype)
        if op == 200: return ~np.asarray(x)
        raise SpirvError(f"opcode {op} ({OPNAMES.get(op, '?')}) not implemented")
