"""Tiny PrettyTable (eval.py:144-146 uses PrettyTable(header), add_rows, print)."""


class PrettyTable:
    def __init__(self, field_names=None):
        self.field_names = list(field_names or [])
        self.rows = []

    def add_row(self, row):
        self.rows.append([str(c) for c in row])

    def add_rows(self, rows):
        for r in rows:
            self.add_row(r)

    def get_string(self):
        cols = [self.field_names] + self.rows
        w = [max(len(str(r[i])) for r in cols) for i in range(len(self.field_names))]
        bar = '+' + '+'.join('-' * (x + 2) for x in w) + '+'
        fmt = lambda r: '|' + '|'.join(' ' + str(c).center(x) + ' ' for c, x in zip(r, w)) + '|'
        return '\n'.join([bar, fmt(self.field_names), bar] + [fmt(r) for r in self.rows] + [bar])

    __str__ = get_string
