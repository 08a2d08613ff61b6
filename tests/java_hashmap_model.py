"""TEST INFRASTRUCTURE: an independent Python restatement of the iteration order of java.util.HashMap<String, ?> under
put() of new keys only, tree bins included (OpenJDK 8..21 HashMap: putVal, resize, treeifyBin, TreeNode.treeify,
putTreeVal, split, untreeify, balanceInsertion, rotateLeft/Right, moveRootToFront). The C++ host mirror carries its own
restatement (csrc/host/java_hashmap_order.hpp); the two were written separately and must agree. Neither could be run
against a JVM in this image."""


def utf16(s):
    b = s.encode("utf-16-be", "surrogatepass")
    return [int.from_bytes(b[i:i + 2], "big") for i in range(0, len(b), 2)]


def string_hash(s):
    h = 0
    for u in utf16(s):
        h = (31 * h + u) & 0xFFFFFFFF
    return h


def _signed(x):
    return x - (1 << 32) if x & 0x80000000 else x


class _N:
    __slots__ = ("key", "u", "hash", "next", "prev", "parent", "left", "right", "red", "tree")

    def __init__(self, key):
        self.key, self.u = key, utf16(key)
        h = string_hash(key)
        self.hash = _signed((h ^ (h >> 16)) & 0xFFFFFFFF)
        self.next = self.prev = self.parent = self.left = self.right = None
        self.red = self.tree = False


class JavaHashMapModel:
    def __init__(self):
        self.tab = [None] * 16
        self.size = 0
        self.keys = set()
        self.treeified = False

    # ---- public ----
    def put(self, key):
        if key in self.keys:
            return
        self.keys.add(key)
        x = _N(key)
        n = len(self.tab)
        i = x.hash & (n - 1)
        first = self.tab[i]
        if first is None:
            self.tab[i] = x
        elif first.tree:
            self._put_tree_val(first, x)
        else:
            p, count = first, 0
            while p.next is not None:
                p = p.next
                count += 1
            p.next = x
            if count >= 7:                     # binCount >= TREEIFY_THRESHOLD - 1
                self._treeify_bin(i)
        self.size += 1
        if self.size > len(self.tab) * 3 // 4:
            self._resize()

    def order(self):
        out = []
        for head in self.tab:
            e = head
            while e is not None:
                out.append(e.key)
                e = e.next
        return out

    # ---- internals ----
    @staticmethod
    def _dir(x, p):
        if p.hash > x.hash:
            return -1
        if p.hash < x.hash:
            return 1
        return -1 if x.u <= p.u else 1        # String.compareTo: UTF-16 units, then length (list comparison does both)

    def _treeify_bin(self, i):
        if len(self.tab) < 64:
            self._resize()
            return
        prev, e = None, self.tab[i]
        while e is not None:
            e.tree, e.prev = True, prev
            e.parent = e.left = e.right = None
            e.red = False
            prev, e = e, e.next
        self._treeify(self.tab[i])

    def _treeify(self, head):
        self.treeified = True
        root, x = None, head
        while x is not None:
            nxt = x.next
            x.left = x.right = None
            if root is None:
                x.parent, x.red, root = None, False, x
            else:
                p = root
                while True:
                    d = self._dir(x, p)
                    xp = p
                    p = p.left if d <= 0 else p.right
                    if p is None:
                        x.parent = xp
                        if d <= 0:
                            xp.left = x
                        else:
                            xp.right = x
                        root = self._balance(root, x)
                        break
            x = nxt
        self._root_to_front(root)

    def _put_tree_val(self, first, x):
        root = first
        while root.parent is not None:
            root = root.parent
        x.tree = True
        p = root
        while True:
            d = self._dir(x, p)
            xp = p
            p = p.left if d <= 0 else p.right
            if p is None:
                xpn = xp.next
                x.next = xpn
                if d <= 0:
                    xp.left = x
                else:
                    xp.right = x
                xp.next = x
                x.parent = x.prev = xp
                if xpn is not None:
                    xpn.prev = x
                self._root_to_front(self._balance(root, x))
                return

    def _root_to_front(self, root):
        i = root.hash & (len(self.tab) - 1)
        first = self.tab[i]
        if root is not first:
            self.tab[i] = root
            rp, rn = root.prev, root.next
            if rn is not None:
                rn.prev = rp
            if rp is not None:
                rp.next = rn
            if first is not None:
                first.prev = root
            root.next, root.prev = first, None

    @staticmethod
    def _rot_left(root, p):
        r = p.right if p is not None else None
        if r is not None:
            p.right = r.left
            if r.left is not None:
                r.left.parent = p
            r.parent = p.parent
            if p.parent is None:
                root = r
                r.red = False
            elif p.parent.left is p:
                p.parent.left = r
            else:
                p.parent.right = r
            r.left, p.parent = p, r
        return root

    @staticmethod
    def _rot_right(root, p):
        l = p.left if p is not None else None
        if l is not None:
            p.left = l.right
            if l.right is not None:
                l.right.parent = p
            l.parent = p.parent
            if p.parent is None:
                root = l
                l.red = False
            elif p.parent.right is p:
                p.parent.right = l
            else:
                p.parent.left = l
            l.right, p.parent = p, l
        return root

    def _balance(self, root, x):
        x.red = True
        while True:
            xp = x.parent
            if xp is None:
                x.red = False
                return x
            if not xp.red or xp.parent is None:
                return root
            xpp = xp.parent
            if xp is xpp.left:
                uncle = xpp.right
                if uncle is not None and uncle.red:
                    uncle.red = xp.red = False
                    xpp.red = True
                    x = xpp
                    continue
                if x is xp.right:
                    x = xp
                    root = self._rot_left(root, x)
                    xp = x.parent
                    xpp = xp.parent if xp is not None else None
                if xp is not None:
                    xp.red = False
                    if xpp is not None:
                        xpp.red = True
                        root = self._rot_right(root, xpp)
            else:
                uncle = xpp.left
                if uncle is not None and uncle.red:
                    uncle.red = xp.red = False
                    xpp.red = True
                    x = xpp
                    continue
                if x is xp.left:
                    x = xp
                    root = self._rot_right(root, x)
                    xp = x.parent
                    xpp = xp.parent if xp is not None else None
                if xp is not None:
                    xp.red = False
                    if xpp is not None:
                        xpp.red = True
                        root = self._rot_left(root, xpp)

    def _resize(self):
        old, oc = self.tab, len(self.tab)
        self.tab = [None] * (2 * oc)
        for j, head in enumerate(old):
            if head is None:
                continue
            was_tree = head.tree
            lists = {0: [], 1: []}
            e = head
            while e is not None:
                nxt = e.next
                lists[1 if e.hash & oc else 0].append(e)
                e = nxt
            for side, idx in ((0, j), (1, j + oc)):
                lst = lists[side]
                for k, e in enumerate(lst):
                    e.prev = lst[k - 1] if k else None
                    e.next = lst[k + 1] if k + 1 < len(lst) else None
                self.tab[idx] = lst[0] if lst else None
            if was_tree:
                for side in (0, 1):
                    lst, other = lists[side], lists[1 - side]
                    if not lst:
                        continue
                    if len(lst) <= 6:
                        for e in lst:
                            e.tree = e.red = False
                            e.parent = e.left = e.right = e.prev = None
                    elif other:
                        self._treeify(lst[0])
