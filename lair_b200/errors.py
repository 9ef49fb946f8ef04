"""`InvalidInput` -- the reference's error enum (src/lib.rs:23-30)."""


class InvalidInput(ValueError):
    """An error returned when a function argument is invalid."""


class Shape(InvalidInput):
    """`InvalidInput::Shape(String)`: displayed as "shape error: {0}"."""

    def __str__(self):
        return f"shape error: {self.args[0]}"


class Value(InvalidInput):
    """`InvalidInput::Value(String)`: displayed as "value error: {0}"."""

    def __str__(self):
        return f"value error: {self.args[0]}"


InvalidInput.Shape = Shape
InvalidInput.Value = Value
