"""A minimal `vitest` for tsinterp: describe / it / test / expect / beforeAll / beforeEach, enough to run the reference's
own test files (tests/*.test.ts) against the reference's own sources under the interpreter.  `Math.random` becomes a
seeded generator so that a run is reproducible.  Test infrastructure only."""
import math

import tsinterp as T


class Node:
    def __init__(self, name, parent):
        self.name, self.parent, self.before_all, self.before_each, self.items = name, parent, [], [], []

    def path(self):
        out, n = [], self
        while n is not None and n.parent is not None:
            out.append(n.name)
            n = n.parent
        return " > ".join(reversed(out))


def _fail(msg):
    raise T.JSThrow(T.JSObj(T.ERR["Error"], {"message": msg, "name": "AssertionError", "stack": ""}))


def _deep_equal(a, b):
    if isinstance(a, T.TypedArray):
        a = [float(x) for x in a.a]
    if isinstance(b, T.TypedArray):
        b = [float(x) for x in b.a]
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_deep_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, (dict, T.JSObj)) and isinstance(b, (dict, T.JSObj)):
        ka = {k for k in T.own_keys(a) if T.get_prop(a, k) is not T.UNDEF}
        kb = {k for k in T.own_keys(b) if T.get_prop(b, k) is not T.UNDEF}
        return ka == kb and all(_deep_equal(T.get_prop(a, k), T.get_prop(b, k)) for k in ka)
    if type(a) is float and type(b) is float and a != a and b != b:
        return True
    return T.strict_eq(a, b)


def _same_value(a, b):   # Object.is
    if type(a) is float and type(b) is float:
        if a != a and b != b:
            return True
        return a == b and math.copysign(1.0, a) == math.copysign(1.0, b)
    return T.strict_eq(a, b)


def make_expect(counter):
    def expect(value=T.UNDEF, *_):
        def matchers(negate):
            def check(ok, what):
                counter[0] += 1
                if bool(ok) == negate:
                    _fail(f"expected {T.to_str(value) if not isinstance(value, (dict, list)) else T.json_stringify(value)} "
                          f"{'not ' if negate else ''}{what}")
                return T.UNDEF

            def num(v):
                return T.to_number(v)

            def to_throw(expected=T.UNDEF):
                try:
                    T.call_function(value, T.UNDEF, [])
                except T.JSThrow as e:
                    msg = T.to_str(T.get_prop(e.value, "message")) if isinstance(e.value, T.JSObj) else T.to_str(e.value)
                    ok = expected is T.UNDEF or (isinstance(expected, str) and expected in msg)
                    return check(ok, f"to throw {expected!r} (threw {msg!r})")
                return check(False, "to throw")

            def has_property(name, *val):
                obj_ok = value is not None and value is not T.UNDEF
                cur = value
                if obj_ok:
                    for part in T.to_str(name).split("."):
                        if cur is None or cur is T.UNDEF or not (T._has_prop(cur, part) if isinstance(cur, (dict, T.JSObj, list, T.TypedArray)) else False):
                            obj_ok = False
                            break
                        cur = T.get_prop(cur, part)
                if val:
                    obj_ok = obj_ok and _deep_equal(cur, val[0])
                return check(obj_ok, f"to have property {name}")

            def close_to(expected, digits=2.0):
                a, e = num(value), num(expected)
                ok = (a == e) or abs(a - e) < (10.0 ** (-T.to_number(digits))) / 2.0
                return check(ok, f"to be close to {e}")

            def length_of(v):
                return T.get_prop(v, "length") if v is not None and v is not T.UNDEF else T.UNDEF

            m = {
                "toBe": lambda e=T.UNDEF: check(_same_value(value, e), f"to be {T.to_str(e)}"),
                "toEqual": lambda e=T.UNDEF: check(_deep_equal(value, e), "to equal the expected value"),
                "toStrictEqual": lambda e=T.UNDEF: check(_deep_equal(value, e), "to strictly equal the expected value"),
                "toHaveLength": lambda n: check(T.strict_eq(length_of(value), n), f"to have length {T.to_str(n)}"),
                "toBeGreaterThan": lambda e: check(num(value) > num(e), f"to be greater than {T.to_str(e)}"),
                "toBeGreaterThanOrEqual": lambda e: check(num(value) >= num(e), f"to be >= {T.to_str(e)}"),
                "toBeLessThan": lambda e: check(num(value) < num(e), f"to be less than {T.to_str(e)}"),
                "toBeLessThanOrEqual": lambda e: check(num(value) <= num(e), f"to be <= {T.to_str(e)}"),
                "toBeCloseTo": close_to, "toHaveProperty": has_property, "toThrow": to_throw, "toThrowError": to_throw,
                "toBeDefined": lambda: check(value is not T.UNDEF, "to be defined"),
                "toBeUndefined": lambda: check(value is T.UNDEF, "to be undefined"),
                "toBeNull": lambda: check(value is None, "to be null"),
                "toBeTruthy": lambda: check(T.truthy(value), "to be truthy"),
                "toBeFalsy": lambda: check(not T.truthy(value), "to be falsy"),
                "toBeNaN": lambda: check(type(value) is float and value != value, "to be NaN"),
                "toContain": lambda e: check(any(_same_value(x, e) for x in T.iterate(value)) if not isinstance(value, str) else T.to_str(e) in value,
                                             f"to contain {T.to_str(e)}"),
                "toBeInstanceOf": lambda c: check(T._instanceof(value, c), "to be an instance of the class"),
            }
            return m
        m = matchers(False)
        m["not"] = matchers(True)
        return m
    return expect


class Vitest:
    def __init__(self):
        self.root = Node("", None)
        self.cur = self.root
        self.assertions = [0]

    def exports(self):
        def describe(name, fn=T.UNDEF, *_):
            node = Node(T.to_str(name), self.cur)
            self.cur.items.append(node)
            prev, self.cur = self.cur, node
            try:
                T.call_function(fn, T.UNDEF, [])
            finally:
                self.cur = prev
            return T.UNDEF

        def it(name, fn=T.UNDEF, *_):
            self.cur.items.append((T.to_str(name), fn))
            return T.UNDEF

        def hook(kind):
            def reg(fn=T.UNDEF, *_):
                getattr(self.cur, kind).append(fn)
                return T.UNDEF
            return reg
        skip = lambda *a: T.UNDEF
        d = T.Native("describe", call=describe, props={"skip": skip, "only": describe})
        i = T.Native("it", call=it, props={"skip": skip, "only": it})
        return {"describe": d, "it": i, "test": i, "expect": make_expect(self.assertions), "beforeAll": hook("before_all"),
                "beforeEach": hook("before_each"), "afterAll": skip, "afterEach": skip,
                "vi": T.Native("vi", props={"fn": lambda *a: (lambda *b: T.UNDEF)})}

    def run(self, log=print):
        """-> (passed, failed: [(path, message)])"""
        passed, failed = [], []

        def each_hooks(node):
            chain = []
            while node is not None:
                chain.append(node)
                node = node.parent
            return [h for n in reversed(chain) for h in n.before_each]

        def walk(node):
            try:
                for h in node.before_all:
                    T.call_function(h, T.UNDEF, [])
            except T.JSThrow as e:
                failed.append((node.path() + " [beforeAll]", str(e)))
                return
            for item in node.items:
                if isinstance(item, Node):
                    walk(item)
                    continue
                name, fn = item
                path = (node.path() + " > " if node.path() else "") + name
                try:
                    for h in each_hooks(node):
                        T.call_function(h, T.UNDEF, [])
                    T.call_function(fn, T.UNDEF, [])
                    passed.append(path)
                except T.JSThrow as e:
                    failed.append((path, str(e)))
        walk(self.root)
        return passed, failed


def seeded_random(seed=12345):
    state = [seed & 0xFFFFFFFF or 1]

    def rnd():   # xorshift32 -> [0, 1)
        x = state[0]
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        state[0] = x
        return x / 4294967296.0
    return rnd


def run_test_file(path, reference_stub=("/src/wasm/index.ts",), seed=12345, log=None, **interp_kwargs):
    """Loads one *.test.ts of the reference (its imports resolve to the reference's own src/ — or to whatever
    path_overrides / require in interp_kwargs put there), runs it -> (passed, failed, assertions, console)"""
    console = []
    vt = Vitest()
    interp = T.Interp(log=(log or (lambda *a: console.append(" ".join(map(str, a))))), stub_modules=reference_stub,
                      virtual_modules={"vitest": vt.exports()}, **interp_kwargs)
    interp.globals.vars["Math"].props["random"] = seeded_random(seed)
    interp.load(path)
    passed, failed = vt.run()
    return passed, failed, vt.assertions[0], console
