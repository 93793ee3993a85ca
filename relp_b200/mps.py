"""MPS reader: a restatement of relp's own two-stage importer.

Stage 1 (`parse_fixed` / `parse_free`, reference `src/io/mps/parse/{mod,fixed,free}.rs`): text -> `MPS`, a
structured copy of the file (rows, columns, right-hand sides, ranges, bounds), syntactic checks only plus the
"name not known" / duplicate checks the reference does while reading.  Stage 2 (`MPS.to_general_form`, reference
`src/io/mps/convert.rs`): bounds semantics, ranges, right-hand sides -> the data of a `GeneralForm`.

Behaviour that matters downstream and is easy to get wrong, all taken from the reference:
  * constraint rows are SORTED BY NAME (`check_row_section_consistency`, parse/mod.rs:311-331): a row's index is its
    position in that order, not its position in the file;
  * numbers are exact decimals `[-]digits[.digits]` -> integer / 10^k, no exponents (`number/parse.rs:84-127`);
  * the first N row is the cost row, a second one is an error (parse/mod.rs:281-288); entries of a column in the
    cost row become `cost_values` (:419-426);
  * a negative upper bound does NOT move the lower bound: the implied zero lower bound is filled in only if no other
    bound touched it, as in GLPK (`convert.rs` process_bound / fill_in_default_lower_bounds);
  * RANGES: `b` holds the UPPER end of the row's interval, the relation holds |r| (compute_b, convert.rs:331-395).

Numbers are `fractions.Fraction` (the reference reads `Rational64` and converts to the solve type).
Host-side only; the device never sees text.
"""
from fractions import Fraction

COMMENT_INDICATOR = "*"            # token.rs
NAME = "NAME"
COLUMN_SECTION_MARKER = "'MARKER'"
START_OF_INTEGER = "'INTORG'"
END_OF_INTEGER = "'INTEND'"

SECTIONS = ("ROWS", "COLUMNS", "RHS", "BOUNDS", "RANGES", "ENDATA")

# character ranges of the fields of a fixed-format line (parse/fixed.rs:137-145)
FIELDS = [(0, 1), (1, 3), (4, 12), (14, 22), (24, 36), (39, 47), (49, 61)]


class ParseError(ValueError):
    """io/error.rs `Parse`: the file is syntactically wrong."""


class Inconsistency(ValueError):
    """io/error.rs `Inconsistency`: the file is logically wrong (unknown names, duplicates, ...)."""


def parse_number(text):
    """number/parse.rs:84-127 (`Raw::try_from` + `From<Raw> for Rational64`): optional '-', digits, optional '.'
    and digits.  Anything else (exponents, '+', blanks inside) is a parse error."""
    if text == "":
        raise ParseError("empty number")
    has_minus = text[0] == "-"
    body = text[1:] if has_minus else text

    def part(t, what):
        if t == "":
            return 0
        if not (t.isascii() and t.isdigit()):
            raise ParseError(f'Failed to parse {what} "{t}" as u64.')
        return int(t)

    if "." in body:
        index = body.index(".")
        from_right = len(body) - index - 1
        integer = part(body[:index], "integer part") * 10 ** from_right + part(body[index + 1:], "mantissa part")
    else:
        from_right = 0
        integer = part(body, "entire value")
    value = Fraction(integer, 10 ** from_right)
    return -value if has_minus else value


# ------------------------------------------------------------------------------------------------
# column retrievers (parse/mod.rs:118-160): how the fields of a line are found
# ------------------------------------------------------------------------------------------------
class Free:
    """parse/free.rs: fields are whitespace separated."""

    @staticmethod
    def two_or_three(line_after_name):
        parts = line_after_name.split()
        if not parts:
            raise ParseError("No name found.")
        return parts[0]

    @staticmethod
    def one_and_two(line):
        parts = line.split()
        if len(parts) < 2:
            raise ParseError("Could not read second field" if parts else "Could not read first field")
        return parts[0], parts[1]

    @staticmethod
    def is_column_marker_line(line):
        parts = line.split()
        if len(parts) < 3:
            raise ParseError("Could not read fourth field")
        if parts[1] == COLUMN_SECTION_MARKER:
            return "marker", parts[2], None
        return "data", (parts[0], parts[1], parts[2]), parts[3:]

    @staticmethod
    def two_through_four(line):
        parts = line.split()
        if len(parts) < 3:
            raise ParseError("Could not read fourth field")
        return (parts[0], parts[1], parts[2]), parts[3:]

    @staticmethod
    def five_and_six(rest):
        if len(rest) > 2:
            raise ParseError("Line has more than 6 elements")
        if len(rest) == 2:
            return rest[0], rest[1]
        if len(rest) == 0:
            return None
        raise ParseError("Line has a fifth element, but no sixth")

    @staticmethod
    def one_through_three(line):
        parts = line.split()
        if len(parts) < 3:
            raise ParseError("Could not read third field")
        return (parts[0], parts[1], parts[2]), parts[3:]

    @staticmethod
    def four(rest):
        if not rest:
            raise ParseError("Could not read value for bound.")
        return rest[0]


def _f(line, k):
    return line[FIELDS[k][0]:FIELDS[k][1]]


class Fixed:
    """parse/fixed.rs: fields sit in fixed character ranges (names may contain blanks)."""

    two_or_three = staticmethod(Free.two_or_three)      # :37-42: deferred to the flexible method

    @staticmethod
    def one_and_two(line):
        if len(line) > FIELDS[2][0]:
            name = line[FIELDS[2][0]:min(FIELDS[2][1], len(line))].rstrip()
            if not name:
                raise ParseError("Empty row name.")
            return _f(line, 1), name
        raise ParseError("Line is too short.")

    @staticmethod
    def is_column_marker_line(line):
        if len(line) >= FIELDS[4][1]:
            if _f(line, 3) == COLUMN_SECTION_MARKER:
                if len(line) >= FIELDS[5][1]:
                    return "marker", _f(line, 5), None
                raise ParseError("Line is too short to be a marker line.")
            return "data", (_f(line, 2).rstrip(), _f(line, 3).rstrip(), _f(line, 4).lstrip()), line[FIELDS[4][1]:]
        raise ParseError("Line is too short.")

    @staticmethod
    def two_through_four(line):
        if len(line) >= FIELDS[4][1]:
            return (_f(line, 2).rstrip(), _f(line, 3).rstrip(), _f(line, 4).lstrip()), line[FIELDS[4][1]:]
        raise ParseError("Line is too short.")

    @staticmethod
    def five_and_six(rest):
        off = FIELDS[4][1]
        if len(rest) >= FIELDS[6][1] - off:
            five = rest[FIELDS[5][0] - off:FIELDS[5][1] - off].rstrip()
            six = rest[FIELDS[6][0] - off:FIELDS[6][1] - off].lstrip()
            if five and six:
                return five, six
        return None

    @staticmethod
    def one_through_three(line):
        if len(line) >= FIELDS[3][0]:
            return (_f(line, 1), _f(line, 2).rstrip(), _f(line, 3).rstrip()), line[FIELDS[3][1]:]
        raise ParseError("Line is too short.")

    @staticmethod
    def four(rest):
        end = FIELDS[4][1] - FIELDS[3][1]
        if len(rest) >= end:
            return rest[FIELDS[4][0] - FIELDS[3][1]:end].lstrip()
        raise ParseError("Line doesn't contain another value, it's too short.")


# ------------------------------------------------------------------------------------------------
# stage 1: text -> MPS
# ------------------------------------------------------------------------------------------------
class MPS:
    """io/mps/mod.rs:49-80.  Indices refer to `rows` (sorted by name, cost row excluded) and `columns` (file
    order).  bounds: [(group name, [(column index, (kind, value or None))])] with kind in
    LO UP FX FR MI PL BV LI UI."""

    def __init__(self, name, objective, cost_row_name, cost_values, rows, columns, rhss, ranges, bounds):
        self.name = name
        self.objective = objective              # "minimize" / "maximize"
        self.cost_row_name = cost_row_name
        self.cost_values = cost_values          # [(column index, value)]
        self.rows = rows                        # [(name, "E" | "L" | "G")]
        self.columns = columns                  # [(name, "continuous" | "integer", [(row index, value)])]
        self.rhss = rhss                        # [(group name, [(row index, value)])]
        self.ranges = ranges
        self.bounds = bounds

    def __eq__(self, other):
        return isinstance(other, MPS) and self.__dict__ == other.__dict__

    def __repr__(self):
        return f"MPS({self.__dict__!r})"

    def to_general_form(self):
        return _to_general_form(self)


def into_lines(text):
    """parse/mod.rs:108-114: numbered from 1; comment lines (first non-blank character '*') and EMPTY lines are
    dropped (a line of blanks is kept, as in the reference)."""
    out = []
    for number, line in enumerate(text.splitlines(), start=1):
        if line.lstrip().startswith(COMMENT_INDICATOR) or line == "":
            continue
        out.append((number, line))
    return out


def parse_program_name(location, cr):
    """parse/mod.rs:177-206"""
    if location is None:
        raise ParseError("No line to read, is the file empty?")
    number, line = location
    if len(line) < len(NAME):
        raise ParseError(f"Line too short. (line {number})")
    if line[:len(NAME)] != NAME:
        raise ParseError(f'Expected a "{NAME}" indicator, found "{line[:len(NAME)]}" instead (line {number})')
    return cr.two_or_three(line[len(NAME):])


def _same_section(line):
    return line.startswith(" ")        # parse/mod.rs:767-771


def _next_section(line, acceptable):
    """try_parse_next_section, parse/mod.rs:752-765"""
    if line not in SECTIONS:
        raise ParseError(f'Unknown section header "{line}".')
    if line != "ENDATA" and line not in acceptable:
        raise ParseError(f"Expected one of the {acceptable} section headers, found the {line} section.")
    return line


def _row_type(word):
    """RowType::from_str, parse/mod.rs:818-830: the first character decides"""
    t = word[0:1]
    if t in ("N", "L", "E", "G"):
        return t
    raise ParseError(f'Row type "{word}" unknown.')


def _parse(text, cr):
    """parse/mod.rs:40-94"""
    lines = iter(into_lines(text))
    name = parse_program_name(next(lines, None), cr)

    # OBJSENSE (optional) and the ROWS header: parse/mod.rs:222-262
    first = next(lines, None)
    if first is None:
        raise ParseError("No line to read, is the program more than a name?")
    number, line = first
    head = line.rstrip()
    if head == "ROWS":
        objective = "minimize"
    elif head == "OBJSENSE":
        sense = next(lines, None)
        if sense is None:
            raise ParseError("Program can't end in the OBJSENSE section.")
        rows_line = next(lines, None)
        if rows_line is None or not rows_line[1].startswith("ROWS"):
            raise ParseError("Expected the ROWS section next.")
        s = sense[1].rstrip()
        if s in ("  MINIMIZE", "  MIN"):
            objective = "minimize"
        elif s in ("  MAXIMIZE", "  MAX"):
            objective = "maximize"
        else:
            raise ParseError(f"Can't read objective {s} (line {sense[0]})")
    else:
        raise ParseError(f'Line contents "{line}" were unexpected (line {number})')

    # ROWS: parse/mod.rs:264-309
    cost_row_name = None
    rows = []
    for number, line in lines:
        if _same_section(line):
            rt, rname = cr.one_and_two(line)
            t = _row_type(rt)
            if t == "N":
                if cost_row_name is not None:
                    raise ParseError(f"Second cost row detected. This is not supported. (line {number})")
                cost_row_name = rname
            else:
                rows.append((rname, t))
        else:
            _next_section(line, ("COLUMNS",))
            break
    else:
        raise ParseError("Section ended sooner than expected.")

    # check_row_section_consistency, parse/mod.rs:311-331: rows sorted by name
    if cost_row_name is None:
        raise Inconsistency("No cost name read.")
    rows.sort(key=lambda r: r[0])
    if any(r[0] == cost_row_name for r in rows):
        raise Inconsistency("Cost row name found in other rows.")
    for a, b in zip(rows, rows[1:]):
        if a[0] == b[0]:
            raise Inconsistency(f"Duplicate row name {a[0]} found.")
    row_index = {r[0]: i for i, r in enumerate(rows)}

    # COLUMNS: parse/mod.rs:367-458
    columns, cost_values = [], []
    pending = [None, []]          # column being read: name, (row index, value) pairs
    vtype = "continuous"

    def save_column(new_name):
        if pending[0] is not None:
            values = sorted(pending[1], key=lambda t: t[0])
            if any(a[0] == b[0] for a, b in zip(values, values[1:])):
                raise Inconsistency(f'Duplicate row for column "{pending[0]}"')
            columns.append((pending[0], vtype, values))
        pending[0], pending[1] = new_name, []

    next_section = None
    for number, line in lines:
        if _same_section(line):
            kind, content, rest = cr.is_column_marker_line(line)
            if kind == "marker":
                save_column(None)
                if content == START_OF_INTEGER:
                    vtype = "integer"
                elif content == END_OF_INTEGER:
                    vtype = "continuous"
                else:
                    raise ParseError(f'Marker type "{content}" unknown. (line {number})')
                continue
            cname, rname, vtext = content
            if pending[0] is not None:
                if pending[0] != cname:
                    save_column(cname)
            else:
                pending[0] = cname

            def save_pair(rname, vtext):
                value = parse_number(vtext)
                if rname in row_index:
                    pending[1].append((row_index[rname], value))
                elif rname == cost_row_name:
                    cost_values.append((len(columns), value))
                else:
                    raise Inconsistency(f'Row "{rname}" not known.')

            save_pair(rname, vtext)
            more = cr.five_and_six(rest)
            if more is not None:
                save_pair(*more)
        else:
            next_section = _next_section(line, ("RHS", "RANGES", "BOUNDS"))
            save_column(None)
            break
    else:
        raise ParseError("Section ended sooner than expected.")
    column_index = {c[0]: j for j, c in enumerate(columns)}

    def optional_section(valid_next):
        """RHS / RANGES: parse/mod.rs:527-571 (groups by name, values sorted by row, duplicates in a group rejected)"""
        collector, group, gname = [], [], [None]

        def save_group(new_name):
            if gname[0] is not None:
                values = sorted(group, key=lambda t: t[0])
                for a, b in zip(values, values[1:]):
                    if a[0] == b[0]:
                        raise Inconsistency(f'Duplicate row id "{a[0]}" for group "{gname[0]}"')
                collector.append((gname[0], values))
            gname[0] = new_name
            group.clear()

        for number, line in lines:
            if _same_section(line):
                (g, rname, vtext), rest = cr.two_through_four(line)
                if gname[0] is not None:
                    if gname[0] != g:
                        save_group(g)
                else:
                    gname[0] = g
                pairs = [(rname, vtext)]
                more = cr.five_and_six(rest)
                if more is not None:
                    pairs.append(more)
                for rn, vt in pairs:
                    if rn not in row_index:
                        raise Inconsistency(f'Row "{rn}" not known.')
                    group.append((row_index[rn], parse_number(vt)))
            else:
                nxt = _next_section(line, valid_next)
                save_group(None)
                return collector, nxt
        raise ParseError('Section "COLUMNS" ended sooner than expected.')

    rhss, ranges, bounds = [], [], []
    if next_section == "RHS":
        rhss, next_section = optional_section(("RANGES", "BOUNDS"))
    if next_section == "RANGES":
        ranges, next_section = optional_section(("BOUNDS",))
    seen = set()                                 # check_ranges_consistency, parse/mod.rs:640-649
    for _, values in ranges:
        for i, _v in values:
            if i in seen:
                raise Inconsistency("Each row can have at most one range value")
            seen.add(i)

    if next_section == "BOUNDS":                 # parse/mod.rs:651-748
        group, gname = [], [None]

        def save_bounds(new_name):
            if gname[0] is not None:
                bounds.append((gname[0], sorted(group, key=lambda t: t[0])))    # duplicates allowed here
            gname[0] = new_name
            group.clear()

        for number, line in lines:
            if _same_section(line):
                (btype, bname, cname), rest = cr.one_through_three(line)
                if cname not in column_index:
                    raise Inconsistency(f'Column name "{cname}" unknown')
                if gname[0] is not None:
                    if gname[0] != bname:
                        save_bounds(bname)
                else:
                    gname[0] = bname
                if btype in ("FR", "MI", "PL", "BV"):
                    bt = (btype, None)
                elif btype in ("LO", "UP", "FX", "LI", "UI"):
                    bt = (btype, parse_number(cr.four(rest)))
                elif btype == "SC":
                    raise NotImplementedError("semi-continuous bounds (unimplemented in the reference too)")
                else:
                    raise ParseError(f'Bound type "{btype}" unknown.')
                group.append((column_index[cname], bt))
            else:
                _next_section(line, ())
                save_bounds(None)
                break
        else:
            raise ParseError('Section "COLUMNS" ended sooner than expected.')

    if next(lines, None) is not None:
        raise ParseError("File parsed successfully, but it has nonempty lines at the end")
    return MPS(name, objective, cost_row_name, cost_values, rows, columns, rhss, ranges, bounds)


def parse_fixed(text):
    """io/mps/parse/fixed.rs:29-31"""
    return _parse(text, Fixed)


def parse_free(text):
    """io/mps/parse/free.rs:27-29; `io::mps::parse` (mod.rs:38-42) defaults to this mode"""
    return _parse(text, Free)


parse = parse_free


# ------------------------------------------------------------------------------------------------
# stage 2: MPS -> GeneralForm data (convert.rs)
# ------------------------------------------------------------------------------------------------
class Variable:
    """general_form `Variable` (`ShiftedVariable` in convert.rs:79-92)"""

    def __init__(self, variable_type, cost):
        self.variable_type = variable_type
        self.cost = cost
        self.lower_bound = None
        self.upper_bound = None
        self.shift = Fraction(0)
        self.flipped = False

    def __repr__(self):
        return f"Variable({self.variable_type}, cost={self.cost}, [{self.lower_bound}, {self.upper_bound}])"


class GeneralFormData:
    """The arguments of `GeneralForm::new` (convert.rs:51-59): objective, column-major constraints over the rows in
    the MPS order (sorted by name), one relation per row -- "E", "L", "G" or ("R", |r|) -- `b` (for a ranged row its
    UPPER end), variables, names, fixed cost 0."""

    def __init__(self, objective, columns, nr_rows, constraint_types, b, variables, variable_names, row_names):
        self.objective = objective
        self.columns = columns
        self.nr_rows = nr_rows
        self.constraint_types = constraint_types
        self.b = b
        self.variables = variables
        self.variable_names = variable_names
        self.row_names = row_names
        self.fixed_cost = Fraction(0)


def _replace_existing_with(var, attr, new_value, greater):
    """convert.rs:196-207: keep the tighter of an existing bound and the new one"""
    cur = getattr(var, attr)
    if cur is None or (new_value > cur if greater else new_value < cur):
        setattr(var, attr, new_value)


def process_bounds(variables, bounds):
    """convert.rs:107-194.  Returns nothing; fills lower / upper bounds and variable types."""
    needs_default_lower = [True] * len(variables)
    is_free = [False] * len(variables)
    for _name, values in bounds:
        for j, (kind, value) in values:
            var = variables[j]
            needs_lower, free = False, False
            if kind == "LO":
                _replace_existing_with(var, "lower_bound", value, True)
            elif kind == "UP":
                # the implied zero lower bound gets filled in only if no other lower bound is present (GLPK)
                _replace_existing_with(var, "upper_bound", value, False)
                needs_lower = True
            elif kind == "FX":
                _replace_existing_with(var, "lower_bound", value, True)
                _replace_existing_with(var, "upper_bound", value, False)
            elif kind == "FR":
                if var.lower_bound is not None or var.upper_bound is not None:
                    raise Inconsistency("Variable can't be bounded and free")
                free = True
            elif kind == "MI":
                _replace_existing_with(var, "upper_bound", Fraction(0), False)
            elif kind == "PL":
                _replace_existing_with(var, "lower_bound", Fraction(0), True)
            elif kind == "BV":
                _replace_existing_with(var, "lower_bound", Fraction(0), True)
                _replace_existing_with(var, "upper_bound", Fraction(1), False)
                var.variable_type = "integer"
            elif kind == "LI":
                _replace_existing_with(var, "lower_bound", value, True)
                var.variable_type = "integer"
            elif kind == "UI":
                _replace_existing_with(var, "upper_bound", value, False)
                var.variable_type = "integer"
                needs_lower = True
            else:
                raise NotImplementedError(kind)
            is_free[j] = is_free[j] or free
            needs_default_lower[j] = needs_default_lower[j] and needs_lower
    for j, var in enumerate(variables):
        if is_free[j] and (var.lower_bound is not None or var.upper_bound is not None):
            raise Inconsistency("A variable is both free and bounded.")
    for j, need in enumerate(needs_default_lower):      # fill_in_default_lower_bounds, convert.rs:209-224
        if need:
            variables[j].lower_bound = Fraction(0)


def compute_ranges(rhss, ranges, nr_rows):
    """convert.rs:247-296: one range per row; a ranged row with several right-hand sides needs them equal"""
    if not ranges:
        return []
    range_rows = sorted((t for _n, values in ranges for t in values), key=lambda t: t[0])
    if any(a[0] == b[0] for a, b in zip(range_rows, range_rows[1:])):
        raise Inconsistency("Only one range per row can be specified.")
    seen, duplicates = [False] * nr_rows, []
    for _n, values in rhss:
        for i, _v in values:
            if seen[i]:
                duplicates.append(i)
            else:
                seen[i] = True
    for d in duplicates:
        if any(i == d for i, _ in range_rows):
            vals = [v for _n, values in rhss for i, v in values if i == d]
            if any(v != vals[0] for v in vals):
                raise Inconsistency("Multiple rhs values for a constraint with a range")
    return range_rows


def compute_constraint_types(rows, ranges):
    """convert.rs:298-322: a zero range makes an equality, another one a ranged row"""
    by_row = dict(ranges)
    out = []
    for i, (_name, t) in enumerate(rows):
        if i in by_row:
            out.append("E" if by_row[i] == 0 else ("R", by_row[i]))
        else:
            out.append(t)
    return out


def compute_b(rhss, constraints, rows, nr_rows):
    """convert.rs:331-395.  `constraints` is updated in place (a range becomes its absolute value)."""
    b = [None] * nr_rows
    for i, value in (t for _n, values in rhss for t in values):
        if b[i] is None:
            c = constraints[i]
            if isinstance(c, tuple):
                r = c[1]
                sign = (r > 0) - (r < 0)
                if r < 0:
                    r = -r
                    constraints[i] = ("R", r)
                t = rows[i][1]
                if t == "G":
                    bound = value + r
                elif t == "L":
                    bound = value
                else:   # "E": the sign of r says on which side of b the interval lies
                    bound = value + r if sign >= 0 else value
                b[i] = bound
            else:
                b[i] = value
        else:
            t = rows[i][1]
            if t == "E":
                if value != b[i]:
                    raise Inconsistency(f"Trivial infeasibility: a constraint can't equal both {b[i]} and {value}")
            elif t == "G":
                if value > b[i]:
                    b[i] = value
            else:
                if value < b[i]:
                    b[i] = value
    return [Fraction(0) if v is None else v for v in b]


def _to_general_form(mps):
    """TryInto<GeneralForm> for MPS, convert.rs:29-61"""
    costs = dict(mps.cost_values)
    variables = [Variable(vt, costs.get(j, Fraction(0))) for j, (_n, vt, _v) in enumerate(mps.columns)]
    process_bounds(variables, mps.bounds)
    columns = [list(values) for _n, _t, values in mps.columns]
    names = [n for n, _t, _v in mps.columns]
    nr_rows = len(mps.rows)
    ranges = compute_ranges(mps.rhss, mps.ranges, nr_rows)
    constraint_types = compute_constraint_types(mps.rows, ranges)
    b = compute_b(mps.rhss, constraint_types, mps.rows, nr_rows)
    return GeneralFormData(mps.objective, columns, nr_rows, constraint_types, b, variables, names,
                           [r[0] for r in mps.rows])
