"""stand-in for `inflect` (utils/parse.py:7,10,340-342): regular English plurals and number words up to 99 - enough
for the reference's convert_spec on the fixture prompts; NOT the real package (irregular nouns are not covered)."""

_ONES = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
         "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]


class _Engine:
    def plural_noun(self, noun):
        if noun.endswith(("s", "x", "z", "ch", "sh")):
            return noun + "es"
        if noun.endswith("y") and noun[-2:-1] not in "aeiou":
            return noun[:-1] + "ies"
        return noun + "s"

    def number_to_words(self, n):
        n = int(n)
        if n < 20:
            return _ONES[n]
        if n < 100:
            return _TENS[n // 10] + ("-" + _ONES[n % 10] if n % 10 else "")
        return str(n)


def engine():
    return _Engine()
