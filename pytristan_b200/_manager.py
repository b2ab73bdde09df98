"""Registry object shared by the grid and operator getters.

Mirrors the reference's attribute-bag registry (reference: src/pytristan/_manager.py:1-12):
objects are stored as instance attributes whose *names* are the identifiers, either a single
integer ("7") or a dash-joined composite ("3-1-2", used here for operator keys
grid-axis-order[-accuracy]).
"""


class ObjectManager:
    """Attribute bag; ``nums()`` lists the identifiers of everything stored."""

    def nums(self):
        """Identifiers in insertion order: an ``int`` for a simple key, a ``list`` of ints for
        a composite key (reference: _manager.py:7-12)."""
        out = []
        for name in vars(self):
            parts = [int(tok) for tok in name.split("-") if tok.isdigit()]
            out.append(parts[0] if len(parts) == 1 else parts)
        return out
